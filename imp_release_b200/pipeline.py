"""Image pair -> matches entirely on the GPU ("would let the pipeline go image -> matches on-GPU").

``ImagePairMatcher`` chains the B200 SuperPoint front-end and a matcher (DGNNS / AdaGMN / GM) without a single host
synchronisation: SuperPoint writes fixed-capacity keypoint / score / descriptor tensors plus DEVICE-side keypoint counts
(``SuperPoint.detect_padded``), and the matcher takes those counts through its ``n_keypoints0/1`` data keys -- every kernel
masks by count -- so the host only enqueues launches and several pairs can be in flight.  The reference pipeline
(components/extractors.py -> eval/eval_imp.py) synchronises twice per image (torch.nonzero, .cpu()) and once per pair.
"""
from __future__ import annotations

from typing import Dict

import torch


class ImagePairMatcher:
    def __init__(self, superpoint, matcher, p: float = 0.2):
        if superpoint.config['max_keypoints'] <= 0:
            raise ValueError("ImagePairMatcher needs SuperPoint(max_keypoints > 0): the fixed capacity of the keypoint tensors")
        self.sp, self.matcher, self.p = superpoint, matcher, p

    @classmethod
    def slots(cls, superpoint, matcher, n: int, p: float = 0.2):
        """n pipelines for n pairs in flight on n CUDA streams (one host thread): matcher replicas share the weights and own
        their workspaces; with more than one stream they take the streaming Sinkhorn kernels, because cooperative launches
        (the shared-memory-resident kernel) from several streams can wait for each other's SMs."""
        out = []
        for _ in range(n):
            m = matcher.replica()
            if n > 1:
                m.sinkhorn_resident = False
            out.append(cls(superpoint, m, p))
        return out

    @torch.no_grad()
    def __call__(self, image0: torch.Tensor, image1: torch.Tensor) -> Dict[str, torch.Tensor]:
        """image0 / image1: [1, 1, H, W] grayscale in [0, 1] on the GPU (sizes may differ).  Returns device tensors
        {'keypoints0', 'keypoints1' [K, 2], 'n_keypoints0', 'n_keypoints1' [1] int32, 'indices0' [K] (index into keypoints1
        or -1), 'mscores0' [K]}; entries behind n_keypoints0 are -1 / 0.  Nothing here waits for the GPU."""
        if image0.shape == image1.shape:
            f = self.sp.detect_padded(torch.cat([image0, image1], 0))
            f0 = {k: v[0:1] for k, v in f.items()}
            f1 = {k: v[1:2] for k, v in f.items()}
        else:
            f0, f1 = self.sp.detect_padded(image0), self.sp.detect_padded(image1)
        data = {'image0': image0, 'image1': image1}          # only .shape is read (keypoint normalisation)
        for i, f in ((0, f0), (1, f1)):
            data[f'keypoints{i}'] = f['keypoints']
            data[f'scores{i}'] = f['scores']
            data[f'descriptors{i}'] = f['descriptors']
            data[f'n_keypoints{i}'] = f['n_keypoints']
        out = self.matcher.produce_matches(data, p=self.p, only_last=True)
        return {'keypoints0': f0['keypoints'][0], 'keypoints1': f1['keypoints'][0],
                'n_keypoints0': f0['n_keypoints'], 'n_keypoints1': f1['n_keypoints'],
                'indices0': out['indices0'][-1][0], 'mscores0': out['mscores0'][-1][0]}
