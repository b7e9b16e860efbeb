"""ExtractSuperpoint with the reference's interface (components/extractors.py:50-89): image file -> (kpt [N, 3] = x, y, score in
original-image pixels; desc [N, 256]), on top of the B200 SuperPoint.  Host-side image reading / resizing stays on the host
(cv2), exactly as in the reference."""
from __future__ import annotations

import numpy as np
import torch

from .nets.superpoint import SuperPoint


def resize(img, resize):
    """components/extractors.py:14-24: longest side (one value) or (h, w) (two values)."""
    import cv2
    img_h, img_w = img.shape[0], img.shape[1]
    cur_size = max(img_h, img_w)
    if len(resize) == 1:
        scale1, scale2 = resize[0] / cur_size, resize[0] / cur_size
    else:
        scale1, scale2 = resize[0] / img_h, resize[1] / img_w
    new_h, new_w = int(img_h * scale1), int(img_w * scale2)
    new_img = cv2.resize(img.astype('float32'), (new_w, new_h)).astype('uint8')
    return new_img, np.asarray([scale2, scale1])


class ExtractSuperpoint(object):
    def __init__(self, config):
        default_config = {
            'descriptor_dim': 256,
            'nms_radius': 4,
            'detection_threshold': config['det_th'],   # sic: SuperPoint reads 'keypoint_threshold', so its default applies
            'max_keypoints': config['num_kpt'],
            'remove_borders': 4,
            'weight_path': config.get('weight_path', '../weights/superpoint_v1.pth'),
        }
        self.superpoint_extractor = SuperPoint(default_config)
        self.superpoint_extractor.eval(), self.superpoint_extractor.cuda()
        self.num_kp = config['num_kpt']
        self.padding = config['padding'] if 'padding' in config.keys() else False
        self.resize = config['resize']

    def run_image(self, img: np.ndarray):
        """Grayscale uint8 image [H, W] (already read) -> (kpt [N, 3], desc [N, 256])."""
        scale = 1
        if self.resize[0] != -1:
            img, scale = resize(img, self.resize)
        with torch.no_grad():
            result = self.superpoint_extractor({'image': torch.from_numpy(img / 255.).float()[None, None].cuda()})
        score, kpt, desc = result['scores'][0], result['keypoints'][0], result['descriptors'][0]
        score, kpt, desc = score.cpu().numpy(), kpt.cpu().numpy(), desc.cpu().numpy().T
        kpt = np.concatenate([kpt / scale, score[:, np.newaxis]], axis=-1)
        if self.padding and len(kpt) < self.num_kp:      # components/extractors.py:79-88 (random padding)
            res = int(self.num_kp - len(kpt))
            pad_x, pad_desc = np.random.uniform(size=[res, 2]) * (img.shape[0] + img.shape[1]) / 2, np.random.uniform(size=[res, 256])
            pad_kpt = np.concatenate([pad_x, np.zeros([res, 1])], axis=-1)
            pad_desc = pad_desc / np.linalg.norm(pad_desc, axis=-1)[:, np.newaxis]
            kpt, desc = np.concatenate([kpt, pad_kpt], axis=0), np.concatenate([desc, pad_desc], axis=0)
        return kpt, desc

    def run(self, img_path):
        import cv2
        return self.run_image(cv2.imread(img_path, cv2.IMREAD_GRAYSCALE))
