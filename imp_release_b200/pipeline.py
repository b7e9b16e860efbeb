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


class GraphedImagePairMatcher:
    """The same pipeline as CUDA-graph replay with several pairs in flight: one graph per (slot, image shapes) holds the whole
    image pair -> matches computation (two SuperPoint detections + the matcher's 15 iterations, ~600 launches), so a submit
    costs the host two image copies, one graph launch and a few result clones.  ``ticket = g.submit(img0, img1)``,
    ``out = g.result(ticket)`` (or ``out = g(img0, img1)``); results as from ``ImagePairMatcher``."""

    def __init__(self, superpoint, matcher, slots: int = 4, p: float = 0.2):
        self.device = next(matcher.parameters()).device
        self.slots = [dict(ipm=ipm, stream=torch.cuda.Stream(device=self.device), graphs={})
                      for ipm in ImagePairMatcher.slots(superpoint, matcher, slots, p)]
        self._next = 0
        self.captures = 0

    @staticmethod
    def _reset_caches(m):
        m._sk_cache = {}
        m._last_sk = None
        for k in ('_dist', '_dist_key', '_n_tok_cache'):
            m.__dict__.pop(k, None)

    def _capture(self, slot, shape0, shape1):
        ipm, dev = slot['ipm'], self.device
        m = ipm.matcher
        img0, img1 = torch.zeros(shape0, device=dev), torch.zeros(shape1, device=dev)
        # the graph bakes in the addresses of every cached buffer: start from empty caches and let the graph entry own them
        # afterwards (the model's caches are bounded and evict; see graphed.LatencyMatcher._capture)
        self._reset_caches(m)
        with torch.no_grad():
            for _ in range(2):                       # warm-up on this stream: weight packing, workspaces, kernel attributes
                ipm(img0, img1)
            slot['stream'].synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=slot['stream']):
                out = ipm(img0, img1)
        keep = (m._sk_cache, m.__dict__.get('_dist'), m.__dict__.get('_last_sk'), list(m.engine()._ws.values()),
                dict(ipm.sp._sel_ws))
        self._reset_caches(m)
        self.captures += 1
        return {'graph': g, 'img0': img0, 'img1': img1, 'out': out, 'keep': keep}

    def submit(self, image0: torch.Tensor, image1: torch.Tensor):
        slot = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        cur = torch.cuda.current_stream(self.device)
        s = slot['stream']
        s.wait_stream(cur)                           # the images were produced on the caller's stream
        key = (tuple(image0.shape), tuple(image1.shape))
        with torch.cuda.stream(s):
            e = slot['graphs'].get(key)
            if e is None:
                e = slot['graphs'][key] = self._capture(slot, *key)
            e['img0'].copy_(image0, non_blocking=True)
            e['img1'].copy_(image1, non_blocking=True)
            e['graph'].replay()
            res = {k: v.clone() for k, v in e['out'].items()}
            done = torch.cuda.Event()
            done.record(s)
        for t in res.values():
            t.record_stream(cur)
        res['done'] = done
        return res

    def result(self, ticket):
        torch.cuda.current_stream(self.device).wait_event(ticket['done'])
        return {k: v for k, v in ticket.items() if k != 'done'}

    def __call__(self, image0, image1):
        return self.result(self.submit(image0, image1))
