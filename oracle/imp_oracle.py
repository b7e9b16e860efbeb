"""CPU oracle for the IMP / EIMP matching hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch fp32 restatement (torch-on-CPU used as a numpy-like
array library, no nn.Module, token-major ``[B, N, C]`` tensors) of the reference
algorithm in feixue94/imp-release.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
(``imp_release_b200``) never does and fails loudly without its CUDA library.

Parity pinning: the reference ships no golden vectors or tests, so the
oracle is pinned against the reference itself: ``tests/golden/make_golden.py`` imports the
unmodified reference classes from /root/reference (CPU, with the one-function device patch
of ``sink_algorithm``), runs them on seeded weights/inputs and commits the outputs as
fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against
those fixtures (indices bit-exact, scores <= 1e-5).

Every function cites the reference lines it restates (paths relative to the reference root).
Weights are addressed by the reference's ``state_dict`` keys, so a reference checkpoint can
be fed to the oracle unchanged.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

Tensor = torch.Tensor

# nets/gms.py:17, nets/adgm.py:18 -- which GNN layers reuse the previous attention map.
SHARING_LAYERS = [False, False] * 2 + [False, False, True, True] * 21
NUM_HEADS = 4            # nets/layers.py:157,230 (hard-coded)
SINK_EPS = 1e-8          # nets/layers.py:13
IN_EPS = 1e-3            # nets/layers.py:68

DEFAULT_CONFIG = {       # nets/gm.py:30-44
    'descriptor_dim': 256,
    'keypoint_encoder': [32, 64, 128, 256],
    'GNN_layers': ['self', 'cross'] * 9,
    'sinkhorn_iterations': 20,
    'match_threshold': 0.2,
    'n_layers': 9,
    'n_min_tokens': 256,
    'with_sinkhorn': True,
    'ac_fn': 'relu',
    'norm_fn': 'bn',
}


# Precision-model hook (tests only): when set, every tensor-core operand of the product path
# (projection / attention / score contractions) is passed through it, e.g. TF32 rounding, to
# predict how far the B200 path may drift from this fp32 oracle.  Default: identity.
_operand_round = None


def _r(x: Tensor) -> Tensor:
    return x if _operand_round is None else _operand_round(x)


def tf32_truncate(x: Tensor) -> Tensor:
    """Drop the 13 low mantissa bits (what a tf32 tensor core reads of an fp32 word)."""
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def tf32_round(x: Tensor) -> Tensor:
    """Round-to-nearest (ties away, like cvt.rna.tf32.f32) to 10 mantissa bits."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


# ----------------------------------------------------------------------------- primitives
def normalize_keypoints(kpts: Tensor, image_shape) -> Tensor:
    """nets/layers.py:49-56: centre on the image and divide by 0.7*max(W, H)."""
    _, _, height, width = image_shape
    size = kpts.new_tensor([float(width), float(height)])
    centre = size / 2
    scale = size.max() * 0.7
    return (kpts - centre) / scale


def linear(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """Conv1d(kernel_size=1) on channels-first data == per-token affine map.
    x [B,N,Cin], w [Cout,Cin,1] (reference layout), b [Cout]."""
    return _r(x) @ _r(w[:, :, 0]).t() + b


def instance_norm_tokens(x: Tensor) -> Tensor:
    """InstanceNorm1d(C, eps=1e-3, affine=False, track_running_stats=False)
    (nets/layers.py:68): per sample, per channel, biased variance over the N tokens."""
    mean = x.mean(dim=1, keepdim=True)
    var = x.var(dim=1, unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + IN_EPS)


def mlp(sd: Dict[str, Tensor], prefix: str, idxs: List[int], x: Tensor) -> Tensor:
    """MLP() of nets/layers.py:59-77 with norm_fn='in', ac_fn='relu': Conv1d at sequential
    indices ``idxs`` (0,3,6,...), IN+ReLU after all but the last."""
    for k, i in enumerate(idxs):
        x = linear(x, sd[f'{prefix}.{i}.weight'], sd[f'{prefix}.{i}.bias'])
        if k < len(idxs) - 1:
            x = torch.relu(instance_norm_tokens(x))
    return x


def keypoint_encoder(sd: Dict[str, Tensor], norm_kpts: Tensor, scores: Tensor) -> Tensor:
    """KeypointEncoder.forward (nets/layers.py:80-90): MLP([3,32,64,128,256,256]) on
    (x, y, score).  Returns [B,N,256]."""
    x = torch.cat([norm_kpts, scores[..., None]], dim=-1)
    return mlp(sd, 'kenc.encoder', [0, 3, 6, 9, 12], x)


def split_heads(x: Tensor) -> Tensor:
    """nets/layers.py:119: ``view(B, dim=64, heads=4, N)`` on channels-first data, i.e.
    channel c belongs to head c % 4 at head-dim c // 4.  [B,N,256] -> [B,4,N,64]."""
    b, n, c = x.shape
    return x.view(b, n, c // NUM_HEADS, NUM_HEADS).permute(0, 3, 1, 2)


def merge_heads(x: Tensor) -> Tensor:
    """Inverse of split_heads (nets/layers.py:134 ``view(B, dim*heads, N)``)."""
    b, h, n, d = x.shape
    return x.permute(0, 2, 3, 1).reshape(b, n, d * h)


def attention_probs(q: Tensor, k: Tensor, key_keep: Optional[Tensor]) -> Tensor:
    """nets/layers.py:121-129: softmax(QK^T / sqrt(64)) per head; keys outside ``key_keep``
    ([B,Ns] bool) are filled with -FLT_MAX before the softmax (-> exactly 0)."""
    s = torch.einsum('bhnd,bhmd->bhnm', _r(q), _r(k)) / (q.shape[-1] ** 0.5)
    if key_keep is not None:
        s = s.masked_fill(~key_keep[:, None, None, :], -torch.finfo(s.dtype).max)
    return torch.softmax(s, dim=-1)


def propagation_layer(sd: Dict[str, Tensor], li: int, sharing: bool, x: Tensor, src: Tensor,
                      prob: Optional[Tensor], key_keep: Optional[Tensor],
                      plain_gnn: bool = False) -> Tuple[Tensor, Tensor]:
    """(Shared)AttentionalPropagation.forward (nets/layers.py:139-149, 182-218).
    Returns (delta [B,Nq,256], prob [B,4,Nq,Ns])."""
    p = f'gnn.layers.{li}'
    if not sharing:
        a = f'{p}.attn'
        q = split_heads(linear(x, sd[f'{a}.proj.0.weight'], sd[f'{a}.proj.0.bias']))
        k = split_heads(linear(src, sd[f'{a}.proj.1.weight'], sd[f'{a}.proj.1.bias']))
        v = split_heads(linear(src, sd[f'{a}.proj.2.weight'], sd[f'{a}.proj.2.bias']))
        prob = attention_probs(q, k, key_keep)
        msg = merge_heads(torch.einsum('bhnm,bhmd->bhnd', _r(prob), _r(v)))
        msg = linear(msg, sd[f'{a}.merge.weight'], sd[f'{a}.merge.bias'])
    else:
        v = split_heads(linear(src, sd[f'{p}.proj.weight'], sd[f'{p}.proj.bias']))
        msg = merge_heads(torch.einsum('bhnm,bhmd->bhnd', _r(prob), _r(v)))
        msg = linear(msg, sd[f'{p}.merge.weight'], sd[f'{p}.merge.bias'])
    y = torch.cat([x, msg], dim=-1)
    return mlp(sd, f'{p}.mlp', [0, 3], y), prob


def pad_dustbin(m: Tensor, dustbin: Tensor) -> Tensor:
    """nets/layers.py:39-40: append one column, then one row, filled with ``bin_score``."""
    b, n0, n1 = m.shape
    out = m.new_empty(b, n0 + 1, n1 + 1)
    out[:, :n0, :n1] = m
    out[:, :n0, n1] = dustbin
    out[:, n0, :] = dustbin
    return out


def sink_algorithm(m: Tensor, dustbin: Tensor, iteration: int) -> Tensor:
    """sink_algorithm + sinkhorn (nets/layers.py:27-46): probability-domain matrix scaling of
    p = softmax(M_aug) with marginals r = [1..1, N0+1], c = [1..1, N1+1]; eps=1e-8 in the
    denominators; ends on a column update.  Device-agnostic (the reference hard-codes 'cuda')."""
    ma = pad_dustbin(m, dustbin)
    b, r_n, c_n = ma.shape
    r = ma.new_ones(b, r_n)
    r[:, -1] = r_n
    c = ma.new_ones(b, c_n)
    c[:, -1] = c_n
    p = torch.softmax(ma, dim=-1)
    u = torch.ones_like(r)
    v = torch.ones_like(c)
    for _ in range(iteration):
        u = r / ((p * v[:, None, :]).sum(-1) + SINK_EPS)
        v = c / ((p * u[:, :, None]).sum(-2) + SINK_EPS)
    return p * u[:, :, None] * v[:, None, :]


def dual_softmax(m: Tensor, dustbin: Tensor) -> Tensor:
    """nets/layers.py:20-24."""
    ma = pad_dustbin(m, dustbin)
    return torch.exp(torch.log_softmax(ma, dim=-1) + torch.log_softmax(ma, dim=1))


def compute_matches(scores: Tensor, p: float = 0.2):
    """GM.compute_matches (nets/gm.py:305-320): mutual nearest neighbours on the
    non-dustbin block; ties resolve to the lowest index (CPU torch.max)."""
    inner = scores[:, :-1, :-1]
    max0, max1 = inner.max(2), inner.max(1)
    idx0, idx1 = max0.indices, max1.indices
    ar0 = torch.arange(idx0.shape[1])[None]
    ar1 = torch.arange(idx1.shape[1])[None]
    mutual0 = ar0 == idx1.gather(1, idx0)
    mutual1 = ar1 == idx0.gather(1, idx1)
    zero = scores.new_tensor(0)
    ms0 = torch.where(mutual0, max0.values, zero)
    ms1 = torch.where(mutual1, ms0.gather(1, idx1), zero)
    valid0 = mutual0 & (ms0 > p)
    valid1 = mutual1 & valid0.gather(1, idx1)
    neg = idx0.new_tensor(-1)
    return torch.where(valid0, idx0, neg), torch.where(valid1, idx1, neg), ms0, ms1


def received_attention(prob: Tensor) -> Tensor:
    """nets/adgm.py:424-432 / 557-565: attention received per source token, summed over
    heads and queries and normalised to sum 1.  prob [B,4,Nq,Ns] -> [B,Ns]."""
    s = prob.sum(dim=1).sum(dim=1)
    return s / s.sum(dim=1, keepdim=True)


def pool_select(row_mass: Tensor, a_self: Tensor, a_cross: Tensor, thresh: float) -> Optional[Tensor]:
    """Keep rule of nets/adgm.py:476-485 / 576-585 for one image of one pair.
    row_mass [n] = sum of the Sinkhorn scores over the non-dustbin block, a_* [n] normalised
    received attention.  Returns sorted kept positions (into the current subset) or None when no
    row passes the threshold.  torch.median is the LOWER median."""
    pids = torch.where(row_mass >= thresh)[0]
    if pids.numel() == 0:
        return None
    md_s = torch.median(a_self[pids])
    md_c = torch.median(a_cross[pids])
    aug_s = torch.where(a_self >= md_s)[0]
    aug_c = torch.where(a_cross >= md_c)[0]
    return torch.unique(torch.cat([pids, aug_s, aug_c]))


# ----------------------------------------------------------------------------- the matcher
class Oracle:
    """Functional mirror of GM / DGNNS / AdaGMN (nets/gm.py, nets/gms.py, nets/adgm.py).

    kind: 'GM' | 'DGNNS' | 'AdaGMN'.  ``sd`` is a reference state_dict (fp32 CPU tensors).
    Descriptors cross the per-layer API channels-first ``[1,256,N]`` exactly like the
    reference boundary (eval/matching.py:47-61); internally everything is token-major.
    """

    def __init__(self, kind: str, config: dict, sd: Dict[str, Tensor]):
        assert kind in ('GM', 'DGNNS', 'AdaGMN')
        self.kind = kind
        self.config = {**DEFAULT_CONFIG, **config}
        self.sd = {k: v.detach().to(torch.float32).cpu() for k, v in sd.items()}
        self.n_layers = self.config['n_layers']
        self.names = self.config['GNN_layers']
        self.sinkhorn_iterations = self.config['sinkhorn_iterations']
        self.with_sinkhorn = self.config['with_sinkhorn']
        self.n_min_tokens = self.config['n_min_tokens']
        self.bin_score = self.sd['bin_score']
        self.sharing = SHARING_LAYERS if kind != 'GM' else [False] * len(self.names)
        self.self_prob0 = self.self_prob1 = self.cross_prob0 = self.cross_prob1 = None

    # -- boundary pieces (channels-first in / out, as in eval/matching.py) -------------
    def encode_keypoint(self, norm_kpts0, norm_kpts1, scores0, scores1):
        """GM.encode_keypoint (nets/gm.py:287-288) -> two [B,256,N]."""
        e0 = keypoint_encoder(self.sd, norm_kpts0, scores0).transpose(1, 2)
        e1 = keypoint_encoder(self.sd, norm_kpts1, scores1).transpose(1, 2)
        return e0, e1

    def _layer(self, li, x, src, prob, keep=None):
        return propagation_layer(self.sd, li, self.sharing[li], x, src, prob, keep)

    def forward_one_layer(self, desc0, desc1, M0, M1, layer_i):
        """DGNNS/AdaGMN.forward_one_layer (nets/gms.py:260-282, nets/adgm.py:528-550)."""
        x0, x1 = desc0.transpose(1, 2), desc1.transpose(1, 2)
        if self.names[layer_i] == 'cross':
            d0, self.cross_prob1 = self._layer(layer_i, x0, x1, self.cross_prob1)
            d1, self.cross_prob0 = self._layer(layer_i, x1, x0, self.cross_prob0)
        else:
            d0, self.self_prob0 = self._layer(layer_i, x0, x0, self.self_prob0)
            d1, self.self_prob1 = self._layer(layer_i, x1, x1, self.self_prob1)
        return (x0 + d0).transpose(1, 2), (x1 + d1).transpose(1, 2)

    def _distance_tm(self, x0, x1, layer_id):
        w, b = self.sd[f'final_proj.{layer_id}.weight'], self.sd[f'final_proj.{layer_id}.bias']
        y0, y1 = linear(x0, w, b), linear(x1, w, b)
        return torch.einsum('bnd,bmd->bnm', _r(y0), _r(y1)) / (self.config['descriptor_dim'] ** 0.5)

    def compute_distance(self, desc0, desc1, layer_id=-1):
        """GM.compute_distance (nets/gm.py:290-295)."""
        lid = layer_id % self.n_layers
        return self._distance_tm(desc0.transpose(1, 2), desc1.transpose(1, 2), lid)

    def compute_score(self, dist, dustbin, iteration):
        """GM.compute_score (nets/gm.py:297-303)."""
        if self.with_sinkhorn:
            return sink_algorithm(dist, dustbin, iteration)
        return dual_softmax(dist, dustbin)

    def compute_matches(self, scores, p=0.2):
        return compute_matches(scores, p)

    def pool(self, pred_score, prob00, prob01, prob11, prob10, mscore_th=0.1,
             uncertainty_ratio=1.0, n_min_tokens=256):
        """AdaGMN.pool (nets/adgm.py:552-605), B=1; DGNNS.pool returns (None, None)
        (nets/gms.py:316)."""
        if self.kind != 'AdaGMN':
            return None, None
        n0, n1 = pred_score.shape[1], pred_score.shape[2]   # NB: augmented sizes, as in the reference
        th = mscore_th * uncertainty_ratio
        ids0 = ids1 = None
        if not (n_min_tokens > 0 and n0 <= n_min_tokens):
            ids0 = pool_select(pred_score[0, :-1, :-1].sum(-1), received_attention(prob00)[0],
                               received_attention(prob01)[0], th)
        if not (n_min_tokens > 0 and n1 <= n_min_tokens):
            ids1 = pool_select(pred_score[0, :-1, :-1].sum(0), received_attention(prob10)[0],
                               received_attention(prob11)[0], th)
        return ids0, ids1

    # -- whole-model entry points ------------------------------------------------------
    def _prepare(self, data):
        """Common head of produce_matches (nets/gms.py:141-172): normalise, encode, add."""
        if 'norm_keypoints0' in data and 'norm_keypoints1' in data:
            nk0, nk1 = data['norm_keypoints0'], data['norm_keypoints1']
        elif 'image0' in data and 'image1' in data:
            nk0 = normalize_keypoints(data['keypoints0'], data['image0'].shape)
            nk1 = normalize_keypoints(data['keypoints1'], data['image1'].shape)
        else:
            raise ValueError('Require image shape for keypoint coordinate normalization')
        x0 = data['descriptors0'] + keypoint_encoder(self.sd, nk0, data['scores0'])
        x1 = data['descriptors1'] + keypoint_encoder(self.sd, nk1, data['scores1'])
        return x0, x1

    def _score_and_match(self, x0, x1, ni, p):
        dist = self._distance_tm(x0, x1, ni)
        score = self.compute_score(dist, self.bin_score, self.sinkhorn_iterations)
        i0, i1, m0, m1 = compute_matches(score, p)
        return score, i0, m0

    def produce_matches(self, data, p=0.2, only_last=False, mscore_th=0.1, uncertainty_ratio=1.0):
        if self.kind == 'GM':
            return self._produce_gm(data, p, only_last)
        if self.kind == 'DGNNS':
            return self._produce_dgnns(data, p, only_last)
        return self._produce_adagmn(data, p, mscore_th, uncertainty_ratio)

    def forward(self, data):
        """GM.forward in eval mode, mode=0 (nets/gm.py:252-258)."""
        return self.produce_matches(data)

    def _produce_gm(self, data, p, only_last):
        """GM.produce_matches + AttentionalGNN.forward (nets/gm.py:145-247, nets/layers.py:152-179)."""
        x0, x1 = self._prepare(data)
        outs = []
        for li, name in enumerate(self.names):
            s0, s1 = (x1, x0) if name == 'cross' else (x0, x1)
            d0, _ = self._layer(li, x0, s0, None)
            d1, _ = self._layer(li, x1, s1, None)
            x0, x1 = x0 + d0, x1 + d1
            if name == 'cross':
                outs.append((x0, x1))
        its = [len(outs) - 1] if only_last else range(len(outs))
        res = {'scores': [], 'indices0': [], 'mscores0': []}
        for ni in its:
            lid = (self.n_layers - 1) if only_last else ni
            s, i0, m0 = self._score_and_match(outs[ni][0], outs[ni][1], lid, p)
            res['scores'].append(s); res['indices0'].append(i0); res['mscores0'].append(m0)
        return res

    def _produce_dgnns(self, data, p, only_last):
        """DGNNS.produce_matches (nets/gms.py:139-258)."""
        x0, x1 = self._prepare(data)
        p00 = p11 = p10 = p01 = None
        res = {'indices0': [], 'mscores0': []}
        for ni in range(self.n_layers):
            d0, p00 = self._layer(2 * ni, x0, x0, p00)
            d1, p11 = self._layer(2 * ni, x1, x1, p11)
            x0, x1 = x0 + d0, x1 + d1
            d0, p10 = self._layer(2 * ni + 1, x0, x1, p10)
            d1, p01 = self._layer(2 * ni + 1, x1, x0, p01)
            x0, x1 = x0 + d0, x1 + d1
            if only_last and ni != self.n_layers - 1:
                continue
            _, i0, m0 = self._score_and_match(x0, x1, ni, p)
            res['indices0'].append(i0); res['mscores0'].append(m0)
        return res

    def _produce_adagmn(self, data, p, mscore_th, uncertainty_ratio):
        """AdaGMN.produce_matches (nets/adgm.py:327-526): keys restricted to the kept ids,
        queries never dropped; per-sample Sinkhorn on the kept subsets from ni >= 2; keep
        sets rebuilt at the end of sharing iterations."""
        x0, x1 = self._prepare(data)
        nb, n0, n1 = x0.shape[0], x0.shape[1], x1.shape[1]
        gids0 = [torch.arange(n0) for _ in range(nb)]
        gids1 = [torch.arange(n1) for _ in range(nb)]
        keep0 = keep1 = None                      # [B,N] bool key masks (M00/M01 <- keep0, M11/M10 <- keep1)
        p00 = p11 = p10 = p01 = None
        res = {'scores': None, 'indices0': [], 'mscores0': [], 'kept0': [], 'kept1': []}
        th = mscore_th * uncertainty_ratio
        for ni in range(self.n_layers):
            d0, p00 = self._layer(2 * ni, x0, x0, p00, keep0)
            d1, p11 = self._layer(2 * ni, x1, x1, p11, keep1)
            x0, x1 = x0 + d0, x1 + d1
            d0, p10 = self._layer(2 * ni + 1, x0, x1, p10, keep1)
            d1, p01 = self._layer(2 * ni + 1, x1, x0, p01, keep0)
            x0, x1 = x0 + d0, x1 + d1
            w, b = self.sd[f'final_proj.{ni}.weight'], self.sd[f'final_proj.{ni}.bias']
            y0, y1 = linear(x0, w, b), linear(x1, w, b)
            if ni < 2:                                           # first_it_to_update
                dist = torch.einsum('bnd,bmd->bnm', _r(y0), _r(y1)) / 16.0
                score = self.compute_score(dist, self.bin_score, self.sinkhorn_iterations)
                i0, _, m0, _ = compute_matches(score, p)
                res['indices0'].append(i0); res['mscores0'].append(m0); res['scores'] = score
                continue
            bi0 = torch.full((nb, n0), -1, dtype=torch.long)
            bm0 = torch.zeros(nb, n0)
            update = SHARING_LAYERS[2 * ni]
            if update:
                a00, a01 = received_attention(p00), received_attention(p01)
                a10, a11 = received_attention(p10), received_attention(p11)
                keep0 = torch.zeros(nb, n0, dtype=torch.bool)
                keep1 = torch.zeros(nb, n1, dtype=torch.bool)
            for b_ in range(nb):
                g0, g1 = gids0[b_], gids1[b_]
                dist = (_r(y0[b_, g0]) @ _r(y1[b_, g1]).t())[None] / 16.0
                score = self.compute_score(dist, self.bin_score, self.sinkhorn_iterations)
                i0, _, m0, _ = compute_matches(score, p)
                i0, m0 = i0[0], m0[0]
                valid = i0 >= 0
                bi0[b_, g0[valid]] = g1[i0[valid]]
                bm0[b_, g0] = m0
                res['scores'] = score
                if update:
                    if not (self.n_min_tokens > 0 and g0.numel() <= self.n_min_tokens):
                        sel = pool_select(score[0, :-1, :-1].sum(-1), a00[b_][g0], a01[b_][g0], th)
                        if sel is not None:
                            g0 = g0[sel]
                    if not (self.n_min_tokens > 0 and g1.numel() <= self.n_min_tokens):
                        sel = pool_select(score[0, :-1, :-1].sum(0), a10[b_][g1], a11[b_][g1], th)
                        if sel is not None:
                            g1 = g1[sel]
                    gids0[b_], gids1[b_] = g0, g1
                    keep0[b_, g0] = True
                    keep1[b_, g1] = True
            res['indices0'].append(bi0); res['mscores0'].append(bm0)
            res['kept0'].append([g.numel() for g in gids0]); res['kept1'].append([g.numel() for g in gids1])
        res['scores'] = [res['scores']]
        return res


def attention_flops(n0: int, n1: int, n_layers: int) -> float:
    """QK^T + P.V FLOPs of one pair for DGNNS/AdaGMN: self n0^2+n1^2, cross
    2*n0*n1 per iteration; non-sharing layers do both contractions (4*D), sharing only P.V (2*D)."""
    d = 256
    total = 0.0
    for ni in range(n_layers):
        per = 4 * d if not SHARING_LAYERS[2 * ni] else 2 * d
        total += per * (n0 * n0 + n1 * n1 + 2 * n0 * n1)
    return total
