"""Feature extraction with the reference's interface (components/extractors.py:50-89, ``ExtractSuperpoint``): an image file
goes in, ``(kpt [N, 3] = x, y, score in original-image pixels, desc [N, 256])`` comes out -- on top of the B200 SuperPoint
(imp_release_b200/nets/superpoint.py).  Reading and resizing the image stay on the host (cv2), as in the reference.

Behaviour kept on purpose: the config key ``det_th`` travels as ``detection_threshold``, which SuperPoint never reads (its
default ``keypoint_threshold`` applies); ``padding`` tops the result up to ``num_kpt`` with random keypoints of score 0 and
random unit descriptors.
"""
from __future__ import annotations

import numpy as np
import torch

from .nets.superpoint import SuperPoint


def resize(img, resize):
    """components/extractors.py:14-24.  ``resize`` = [longest side] or [height, width]; returns the uint8 image and the
    (x, y) scale factors that map original to resized coordinates."""
    import cv2
    h, w = img.shape[:2]
    if len(resize) == 1:
        sy = sx = resize[0] / max(h, w)
    else:
        sy, sx = resize[0] / h, resize[1] / w
    out = cv2.resize(img.astype('float32'), (int(w * sx), int(h * sy))).astype('uint8')
    return out, np.asarray([sx, sy])


class ExtractSuperpoint(object):
    def __init__(self, config):
        sp_config = dict(descriptor_dim=256, nms_radius=4, remove_borders=4,
                         detection_threshold=config['det_th'],
                         max_keypoints=config['num_kpt'],
                         weight_path=config.get('weight_path', '../weights/superpoint_v1.pth'))
        self.superpoint_extractor = SuperPoint(sp_config).eval().cuda()
        self.num_kp = config['num_kpt']
        self.padding = bool(config.get('padding', False))
        self.resize = config['resize']

    def _pad(self, kpt, desc, extent):
        """Random filler up to num_kp (components/extractors.py:79-88): positions uniform in [0, extent), score 0."""
        missing = int(self.num_kp - len(kpt))
        if missing <= 0:
            return kpt, desc
        xy = np.random.uniform(size=[missing, 2]) * extent
        filler_desc = np.random.uniform(size=[missing, 256])
        filler_desc /= np.linalg.norm(filler_desc, axis=-1, keepdims=True)
        filler_kpt = np.concatenate([xy, np.zeros([missing, 1])], axis=-1)
        return np.concatenate([kpt, filler_kpt], axis=0), np.concatenate([desc, filler_desc], axis=0)

    def run_image(self, img: np.ndarray):
        """Grayscale uint8 image [H, W] (already decoded) -> (kpt [N, 3], desc [N, 256])."""
        scale = 1
        if self.resize[0] != -1:
            img, scale = resize(img, self.resize)
        image = torch.from_numpy(img / 255.).float()[None, None].cuda()
        with torch.no_grad():
            feats = self.superpoint_extractor({'image': image})
        xy = feats['keypoints'][0].cpu().numpy() / scale
        score = feats['scores'][0].cpu().numpy()
        desc = feats['descriptors'][0].t().contiguous().cpu().numpy()
        kpt = np.concatenate([xy, score[:, None]], axis=-1)
        if self.padding:
            kpt, desc = self._pad(kpt, desc, (img.shape[0] + img.shape[1]) / 2)
        return kpt, desc

    def run(self, img_path):
        import cv2
        return self.run_image(cv2.imread(img_path, cv2.IMREAD_GRAYSCALE))
