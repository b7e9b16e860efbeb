"""Golden fixtures of the SuperPoint front-end, made by the UNMODIFIED reference class (nets/superpoint.py) on CPU.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_superpoint.py
Weights and images come from seeds (oracle/superpoint_oracle.py); the reference insists on a weight file, so the seeded
state dict is written to a temporary .pth.  Stored per case: keypoints, scores, the first 48 descriptors, every descriptor's
projection on 4 fixed directions, and a coarse copy of the dense score map.
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import superpoint_oracle as spo  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'

CASES = {
    # name: (weight seed, image seed, H, W, batch, config overrides)
    'all_120x160': (11, 21, 120, 160, 1, {}),
    'top300_240x320': (11, 22, 240, 320, 1, {'max_keypoints': 300}),
    'ragged_100x150_b2': (12, 23, 100, 150, 2, {'max_keypoints': 150, 'nms_radius': 3}),
}
DEFAULT = {'descriptor_dim': 256, 'nms_radius': 4, 'keypoint_threshold': 0.0025, 'max_keypoints': -1, 'remove_borders': 4}


def probe_dirs():
    return torch.randn(256, 4, generator=torch.Generator().manual_seed(99))


def main():
    sys.path.insert(0, REF)
    from nets.superpoint import SuperPoint  # the reference class
    res = {}
    for name, (wseed, iseed, H, W, B, over) in CASES.items():
        sd = spo.make_state_dict(wseed)
        img = spo.make_image(iseed, H, W, B)
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, 'w.pth')
            torch.save(sd, path)
            net = SuperPoint({**over, 'weight_path': path}).eval()
        with torch.no_grad():
            out = net({'image': img})
            dense_scores, _ = net.extract({'image': img})
        res[f'{name}/weights_checksum'] = np.float64(sum(float(v.double().abs().sum()) for v in sd.values()))
        res[f'{name}/image_checksum'] = np.float64(float(img.double().sum()))
        res[f'{name}/dense_scores_8x'] = dense_scores[:, ::8, ::8].numpy().copy()
        res[f'{name}/dense_scores_sum'] = np.float64(float(dense_scores.double().sum()))
        for b in range(B):
            k, s, d = out['keypoints'][b], out['scores'][b], out['descriptors'][b]
            res[f'{name}/{b}/keypoints'] = k.numpy().astype(np.int32)
            res[f'{name}/{b}/scores'] = s.numpy()
            res[f'{name}/{b}/descriptors_head'] = d[:, :48].numpy().copy()
            res[f'{name}/{b}/descriptor_probes'] = (d.t() @ probe_dirs()).numpy()
            print(name, b, 'keypoints', tuple(k.shape), 'score range', float(s.min()), float(s.max()))
    np.savez_compressed(os.path.join(OUT, 'reference_superpoint.npz'), **res)
    print('wrote', os.path.join(OUT, 'reference_superpoint.npz'), os.path.getsize(os.path.join(OUT, 'reference_superpoint.npz')), 'bytes')


if __name__ == '__main__':
    main()
