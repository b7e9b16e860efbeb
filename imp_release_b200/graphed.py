"""CUDA-graph replay of a whole matcher forward for fixed shapes (the B = 1 latency path of eval/eval_imp.py:155-173).

A forward is ~250 small launches; at one pair per call the host (Python + ctypes + tensor-map encoding) costs more
than the GPU work.  All kernels of libimp_b200.so take raw pointers to persistent workspaces and enqueue on the
current stream without synchronising, so the whole sequence can be captured once per (B, N0, N1) and replayed:
inputs are copied into static buffers, outputs are read from static tensors."""
from __future__ import annotations

from typing import Dict

import torch

_KEYS = ('descriptors0', 'descriptors1', 'keypoints0', 'keypoints1', 'scores0', 'scores1')


class GraphedMatcher:
    def __init__(self, model, example: Dict[str, torch.Tensor], **call_kwargs):
        self.model = model
        self.kw = call_kwargs
        self.static = {k: example[k].clone() for k in _KEYS}
        for k in ('image0', 'image1', 'norm_keypoints0', 'norm_keypoints1'):
            if k in example:
                self.static[k] = example[k].clone() if k.startswith('norm') else example[k]
        self.shapes = {k: tuple(self.static[k].shape) for k in _KEYS}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():      # warm-up: weight packing, workspaces, func attributes
            for _ in range(2):
                model.produce_matches(self.static, **call_kwargs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = model.produce_matches(self.static, **call_kwargs)

    def __call__(self, data: Dict[str, torch.Tensor]):
        for k in _KEYS:
            if tuple(data[k].shape) != self.shapes[k]:
                raise ValueError(f'GraphedMatcher was captured for {k} of shape {self.shapes[k]}, got {tuple(data[k].shape)}')
            self.static[k].copy_(data[k], non_blocking=True)
        for k in ('norm_keypoints0', 'norm_keypoints1'):
            if k in self.static:
                self.static[k].copy_(data[k], non_blocking=True)
        self.graph.replay()
        return self.out
