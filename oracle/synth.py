"""Seeded synthetic weights and inputs shared by the oracle, the tests and bench.py.
TEST / BENCH INFRASTRUCTURE (lives beside the oracle, never imported by the product).

* ``state_dict_spec`` lists the reference's ``state_dict`` keys and shapes for GM / DGNNS / AdaGMN
  (nets/gm.py:58-75, nets/layers.py:80-90,100-107,182-198).  ``make_golden.py``
  proves the spec by loading it into the reference classes with ``strict=True``.
* ``make_state_dict`` fills it from a ``torch.Generator`` in key order with U(-1,1)/sqrt(fan_in)
  (the scale of the reference's default Conv1d init), so weights are reproducible anywhere
  without shipping a 77 MB checkpoint (pretrained weights are not available offline).
* ``make_pair_batch`` builds the SURVEY.md 8(d) inputs: unit-norm descriptors, image 1 a noisy
  permutation of image 0 so that a useful number of mutual matches exists.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch

SHARING_LAYERS = [False, False] * 2 + [False, False, True, True] * 21


def state_dict_spec(kind: str, n_layers: int, d: int = 256, kenc=(32, 64, 128, 256)) -> "OrderedDict[str, tuple]":
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    spec['bin_score'] = ()
    chans = [3] + list(kenc) + [d]
    for i in range(1, len(chans)):
        spec[f'kenc.encoder.{3 * (i - 1)}.weight'] = (chans[i], chans[i - 1], 1)
        spec[f'kenc.encoder.{3 * (i - 1)}.bias'] = (chans[i],)
    for li in range(2 * n_layers):
        p = f'gnn.layers.{li}'
        sharing = SHARING_LAYERS[li] if kind != 'GM' else False
        if not sharing:
            spec[f'{p}.attn.merge.weight'] = (d, d, 1)
            spec[f'{p}.attn.merge.bias'] = (d,)
            for j in range(3):
                spec[f'{p}.attn.proj.{j}.weight'] = (d, d, 1)
                spec[f'{p}.attn.proj.{j}.bias'] = (d,)
        else:
            spec[f'{p}.proj.weight'] = (d, d, 1)
            spec[f'{p}.proj.bias'] = (d,)
            spec[f'{p}.merge.weight'] = (d, d, 1)
            spec[f'{p}.merge.bias'] = (d,)
        spec[f'{p}.mlp.0.weight'] = (2 * d, 2 * d, 1)
        spec[f'{p}.mlp.0.bias'] = (2 * d,)
        spec[f'{p}.mlp.3.weight'] = (d, 2 * d, 1)
        spec[f'{p}.mlp.3.bias'] = (d,)
    for i in range(n_layers):
        spec[f'final_proj.{i}.weight'] = (d, d, 1)
        spec[f'final_proj.{i}.bias'] = (d,)
    return spec


def make_state_dict(kind: str, n_layers: int, seed: int, bin_score: float = 1.0,
                    gain: float = 1.0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for k, shape in state_dict_spec(kind, n_layers).items():
        if k == 'bin_score':
            sd[k] = torch.tensor(float(bin_score))
            continue
        fan_in = shape[1] if len(shape) == 3 else None
        if fan_in is None:                      # bias: bound by the matching weight's fan-in
            fan_in = sd[k.replace('.bias', '.weight')].shape[1]
        t = (torch.rand(shape, generator=g) * 2 - 1) * (gain / fan_in ** 0.5)
        sd[k] = t
    return sd


def make_pair_batch(seed: int, batch: int, n0: int, n1: int, width: int = 640, height: int = 480,
                    noise: float = 0.3, d: int = 256) -> Dict[str, torch.Tensor]:
    """Synthetic pair batch.  Image 1 = noisy permutation of (a prefix/extension of) image 0."""
    g = torch.Generator().manual_seed(seed)
    nmax = max(n0, n1)
    base_d = torch.nn.functional.normalize(torch.randn(batch, nmax, d, generator=g), dim=-1)
    base_k = torch.rand(batch, nmax, 2, generator=g) * torch.tensor([float(width), float(height)])
    base_s = torch.rand(batch, nmax, generator=g)
    perm = torch.stack([torch.randperm(nmax, generator=g) for _ in range(batch)])
    d1 = torch.gather(base_d, 1, perm[..., None].expand(-1, -1, d))
    d1 = torch.nn.functional.normalize(d1 + noise * torch.randn(batch, nmax, d, generator=g) / 16.0, dim=-1)
    k1 = torch.gather(base_k, 1, perm[..., None].expand(-1, -1, 2)) + 2.0 * torch.randn(batch, nmax, 2, generator=g)
    s1 = torch.gather(base_s, 1, perm)
    return {
        'descriptors0': base_d[:, :n0].contiguous(), 'descriptors1': d1[:, :n1].contiguous(),
        'keypoints0': base_k[:, :n0].contiguous(), 'keypoints1': k1[:, :n1].contiguous(),
        'scores0': base_s[:, :n0].contiguous(), 'scores1': s1[:, :n1].contiguous(),
        'image0': torch.zeros(1, 1, height, width), 'image1': torch.zeros(1, 1, height, width),
    }


def tensor_checksum(t: torch.Tensor) -> float:
    """Order-dependent fp64 checksum used to detect RNG drift of the seeded fixtures."""
    f = t.detach().double().flatten()
    w = torch.arange(1, f.numel() + 1, dtype=torch.float64) % 97 + 1
    return float((f * w).sum())


def state_dict_checksum(sd: Dict[str, torch.Tensor]) -> float:
    return float(sum(tensor_checksum(v) for v in sd.values()))


def make_scene_pair(seed: int, n0: int, n1: int, width: int = 1600, height: int = 1200, noise: float = 0.3, d: int = 256):
    """One synthetic pair WITH epipolar geometry (for the host pose / RANSAC leg): random 3-D
    points seen by two pinhole cameras; keypoints = their projections (+ 0.5 px noise), descriptors = one unit vector per
    3-D point (+ noise in image 1), image 1 in permuted order.  Returns the reference's feed dict (eval/eval_imp.py:59-78)
    incl. 'K0', 'K1', 'T_0to1', 'pts0_cpu', 'pts1_cpu'."""
    import numpy as np
    g = torch.Generator().manual_seed(seed)
    n = max(n0, n1)
    f = 0.9 * width
    K = np.array([[f, 0., width / 2], [0., f, height / 2], [0., 0., 1.]])
    uv0 = torch.rand(n, 2, generator=g) * torch.tensor([float(width), float(height)])
    depth = 4.0 + 4.0 * torch.rand(n, generator=g)
    x0 = torch.stack([(uv0[:, 0] - width / 2) / f * depth, (uv0[:, 1] - height / 2) / f * depth, depth], 1).double()
    ang = 0.15
    R = torch.tensor([[np.cos(ang), 0., np.sin(ang)], [0., 1., 0.], [-np.sin(ang), 0., np.cos(ang)]], dtype=torch.float64)
    t = torch.tensor([-0.8, 0.05, 0.1], dtype=torch.float64)
    x1 = x0 @ R.T + t
    uv1 = torch.stack([x1[:, 0] / x1[:, 2] * f + width / 2, x1[:, 1] / x1[:, 2] * f + height / 2], 1).float()
    uv1 = uv1 + 0.5 * torch.randn(n, 2, generator=g)
    desc = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1)
    perm = torch.randperm(n, generator=g)
    d1 = torch.nn.functional.normalize(desc[perm] + noise * torch.randn(n, d, generator=g) / 16.0, dim=-1)
    s = torch.rand(n, generator=g)
    k0, k1 = uv0[:n0], uv1[perm][:n1]
    return {
        'descriptors0': desc[:n0][None].contiguous(), 'descriptors1': d1[:n1][None].contiguous(),
        'keypoints0': k0[None].contiguous(), 'keypoints1': k1[None].contiguous(),
        'scores0': s[:n0][None].contiguous(), 'scores1': s[perm][:n1][None].contiguous(),
        'image0': torch.zeros(1, 1, height, width), 'image1': torch.zeros(1, 1, height, width),
        'K0': K, 'K1': K.copy(), 'T_0to1': np.hstack([R.numpy(), t.numpy().reshape(3, 1)]),
        'pts0_cpu': k0.numpy(), 'pts1_cpu': k1.numpy(), 'perm': perm,
    }
