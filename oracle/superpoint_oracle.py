"""CPU restatement of the reference SuperPoint front-end (nets/superpoint.py) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path
(imp_release_b200/) never does.  Functional, fp32, torch as the array library.  Pinned against the unmodified reference by
tests/golden/make_golden_superpoint.py -> tests/golden/reference_superpoint.npz (tests/test_oracle_golden.py).

``make_state_dict`` / ``make_image`` give seeded weights and inputs (the published superpoint_v1.pth is not available
offline), reproducible on any box.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

LAYERS = [('conv1a', 1, 64, 3), ('conv1b', 64, 64, 3), ('conv2a', 64, 64, 3), ('conv2b', 64, 64, 3), ('conv3a', 64, 128, 3),
          ('conv3b', 128, 128, 3), ('conv4a', 128, 128, 3), ('conv4b', 128, 128, 3), ('convPa', 128, 256, 3), ('convPb', 256, 65, 1),
          ('convDa', 128, 256, 3), ('convDb', 256, 256, 1)]   # nets/superpoint.py:122-143


def make_state_dict(seed: int) -> "OrderedDict[str, torch.Tensor]":
    """He-scaled normal weights (keeps activations O(1) through the ReLU stack), small biases."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, cin, cout, k in LAYERS:
        sd[f'{name}.weight'] = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
        sd[f'{name}.bias'] = torch.randn(cout, generator=g) * 0.05
    return sd


def make_image(seed: int, H: int, W: int, batch: int = 1) -> torch.Tensor:
    """Smooth random texture in [0, 1]: low-resolution noise upsampled bilinearly + fine noise."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(batch, 1, max(H // 12, 2), max(W // 12, 2), generator=g)
    img = F.interpolate(low, size=(H, W), mode='bilinear', align_corners=False)
    img = 0.75 * img + 0.25 * torch.rand(batch, 1, H, W, generator=g)
    return (img * 255).round() / 255.0     # what `img / 255.` of a uint8 image gives (components/extractors.py:72)


def dense(sd, image):
    """nets/superpoint.py:186-203, 228-230: -> (scores [B, 8Hc, 8Wc] before NMS, descriptors [B, 256, Hc, Wc] normalised)."""
    c = lambda x, n, pad: F.conv2d(x, sd[f'{n}.weight'], sd[f'{n}.bias'], padding=pad)
    x = image
    for i, n in enumerate(('conv1a', 'conv1b', 'conv2a', 'conv2b', 'conv3a', 'conv3b', 'conv4a', 'conv4b')):
        x = F.relu(c(x, n, 1))
        if n in ('conv1b', 'conv2b', 'conv3b'):
            x = F.max_pool2d(x, 2, 2)
    logits = c(F.relu(c(x, 'convPa', 1)), 'convPb', 0)
    p = torch.softmax(logits, 1)[:, :-1]                       # drop the dust bin
    B, _, Hc, Wc = p.shape
    scores = p.permute(0, 2, 3, 1).reshape(B, Hc, Wc, 8, 8).permute(0, 1, 3, 2, 4).reshape(B, Hc * 8, Wc * 8)
    d = c(F.relu(c(x, 'convDa', 1)), 'convDb', 0)
    return scores, F.normalize(d, p=2, dim=1)


def nms(scores, radius):
    """simple_nms, nets/superpoint.py:50-66."""
    mp = lambda t: F.max_pool2d(t, 2 * radius + 1, 1, radius)
    keep = scores == mp(scores)
    for _ in range(2):
        supp = mp(keep.float()) > 0
        rest = torch.where(supp, torch.zeros_like(scores), scores)
        keep = keep | ((rest == mp(rest)) & ~supp)
    return torch.where(keep, scores, torch.zeros_like(scores))


def detect(scores_nms, threshold, border, max_keypoints):
    """One image [H, W]: nets/superpoint.py:206-225 -> (keypoints [K, 2] as (x, y) float, scores [K])."""
    H, W = scores_nms.shape
    yx = torch.nonzero(scores_nms > threshold)
    s = scores_nms[yx[:, 0], yx[:, 1]]
    ok = (yx[:, 0] >= border) & (yx[:, 0] < H - border) & (yx[:, 1] >= border) & (yx[:, 1] < W - border)
    yx, s = yx[ok], s[ok]
    if max_keypoints >= 0 and max_keypoints < len(yx):
        s, idx = torch.topk(s, max_keypoints, dim=0)
        yx = yx[idx]
    return torch.flip(yx, [1]).float(), s


def sample(kpts_xy, dmap, s=8):
    """sample_descriptors, nets/superpoint.py:83-95 (grid_sample runs with its default align_corners=False: the reference's
    version test int(torch.__version__[2]) > 2 is false for '1.12' and '2.x' alike).  dmap [256, Hc, Wc] -> [256, K]."""
    c, h, w = dmap.shape
    k = kpts_xy - s / 2 + 0.5
    k = k / torch.tensor([w * s - s / 2 - 0.5, h * s - s / 2 - 0.5]).to(k)[None]
    k = k * 2 - 1
    d = F.grid_sample(dmap[None], k.view(1, 1, -1, 2), mode='bilinear', align_corners=False)
    return F.normalize(d.reshape(1, c, -1), p=2, dim=1)[0]


def forward(sd, image, config):
    """-> {'keypoints': [...], 'scores': [...], 'descriptors': [...]} like SuperPoint.forward."""
    scores, dmap = dense(sd, image)
    snms = nms(scores, config['nms_radius'])
    out = {'keypoints': [], 'scores': [], 'descriptors': []}
    for b in range(image.shape[0]):
        k, s = detect(snms[b], config['keypoint_threshold'], config['remove_borders'], config['max_keypoints'])
        out['keypoints'].append(k)
        out['scores'].append(s)
        out['descriptors'].append(sample(k, dmap[b]))
    return out
