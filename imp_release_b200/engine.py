"""Host-side execution engine: weight packing, workspaces and the kernel schedule of one matcher forward.

Everything numeric happens in libimp_b200.so (see ops.py); this file only orders launches on the current CUDA
stream.  Layout: the two images of every pair are stacked image-major, ``img = side * B + b`` (side 0 = image 0),
tokens are token-major ``[2B, Np, C]`` with ``Np = max(N0, N1)``; per-image valid counts live in ``n_tok``.

Kernel schedule per GNN layer (Appendix A of SURVEY.md; reference nets/layers.py:200-218, 109-136):
  1. fused Q|K|V projection (one split-precision GEMM, N = 768, heads de-interleaved via permuted weight rows)
     -- sharing layers project V only and reuse the stashed Q, K and row LSE of the previous iteration
  2. flash attention (tcgen05, fp16 operands) -> message in head-contiguous channels, fp16 hi/lo planes
  3. [x | msg] . W0'^T + b0'  with  W0' = [W0a, W0b . Wmerge]  (merge conv folded into the first MLP conv)
  4. instance norm over the tokens of each image + ReLU
  5. . W3^T + b3 + residual -> x (in place)
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops
from .ops import Planes

HEADS = 4
D = 256
# instance-norm statistics fused into the epilogue of the first MLP GEMM (IMP_FUSED_INSTNORM=0: stand-alone slab kernel)
import os as _os
FUSED_INSTNORM = _os.environ.get('IMP_FUSED_INSTNORM', '1') != '0'
# ... and the normalisation itself folded into the A-operand path of the second MLP GEMM (IMP_FUSED_NORM_A=0: apply pass)
FUSED_NORM_A = _os.environ.get('IMP_FUSED_NORM_A', '1') != '0'


def head_perm(device) -> torch.Tensor:
    """perm[j'] = original channel of head-contiguous channel j' = h*64 + d  (orig c = d*4 + h, nets/layers.py:119)."""
    j = torch.arange(D, device=device)
    return (j % 64) * HEADS + j // 64


def _planes_of(w: torch.Tensor) -> Planes:
    return ops.split_planes(w.contiguous().float())


class PackedModel:
    """Device-resident, kernel-ready weights built once from the module's parameters."""

    def __init__(self, sd: Dict[str, torch.Tensor], n_layers: int, sharing: List[bool]):
        dev = sd['bin_score'].device
        perm = head_perm(dev)
        f = lambda k: sd[k].detach().float()
        w2 = lambda k: f(k)[:, :, 0]
        self.kenc_small = [(w2('kenc.encoder.0.weight').contiguous(), f('kenc.encoder.0.bias').contiguous()),
                           (w2('kenc.encoder.3.weight').contiguous(), f('kenc.encoder.3.bias').contiguous())]
        self.kenc_big = [(_planes_of(w2(f'kenc.encoder.{i}.weight')), f(f'kenc.encoder.{i}.bias').contiguous(),
                          w2(f'kenc.encoder.{i}.weight').shape) for i in (6, 9, 12)]
        self.layers = []
        for li in range(2 * n_layers):
            p = f'gnn.layers.{li}'
            L = {'sharing': sharing[li]}
            if not sharing[li]:
                a = f'{p}.attn'
                wq = [w2(f'{a}.proj.{j}.weight')[perm] for j in range(3)]
                bq = [f(f'{a}.proj.{j}.bias')[perm] for j in range(3)]
                L['Wqkv'] = _planes_of(torch.cat(wq, 0))
                L['bqkv'] = torch.cat(bq, 0).contiguous()
                wm, bm = w2(f'{a}.merge.weight'), f(f'{a}.merge.bias')
            else:
                L['Wv'] = _planes_of(w2(f'{p}.proj.weight')[perm])
                L['bv'] = f(f'{p}.proj.bias')[perm].contiguous()
                wm, bm = w2(f'{p}.merge.weight'), f(f'{p}.merge.bias')
            w0, b0 = w2(f'{p}.mlp.0.weight'), f(f'{p}.mlp.0.bias')
            # fold the merge conv: [x, A.Wm^T + bm].W0^T + b0 = x.W0a^T + A.(W0b.Wm)^T + (b0 + W0b.bm)
            w0b = w0[:, D:].double()
            fused = torch.cat([w0[:, :D].double(), w0b @ wm[:, perm].double()], 1).float()
            L['W0'] = _planes_of(fused)
            L['b0'] = (b0.double() + w0b @ bm.double()).float().contiguous()
            L['W3'] = _planes_of(w2(f'{p}.mlp.3.weight'))
            L['b3'] = f(f'{p}.mlp.3.bias').contiguous()
            self.layers.append(L)
        self.final = [(_planes_of(w2(f'final_proj.{i}.weight')), f(f'final_proj.{i}.bias').contiguous())
                      for i in range(n_layers)]


class Workspace:
    """Per-(n_img, Np) device buffers, zero-initialised once and reused across calls."""

    def __init__(self, n_img: int, Np: int, device):
        self.n_img, self.Np = n_img, Np
        T = n_img * Np
        f16 = dict(dtype=torch.float16, device=device)
        f32 = dict(dtype=torch.float32, device=device)
        self.X = Planes.empty((T, D), device)
        self.A = Planes.empty((T, D), device)
        self.Hn = Planes.empty((T, 2 * D), device)
        self.Y = Planes.empty((T, D), device)
        self.H = torch.zeros(T, 2 * D, **f32)
        self.qkv = {'self': torch.zeros(T, 3 * D, **f16), 'cross': torch.zeros(T, 3 * D, **f16)}
        self.lse = {'self': torch.zeros(n_img, HEADS, Np, **f32), 'cross': torch.zeros(n_img, HEADS, Np, **f32)}
        # EIMP: compacted [K | V] rows of the kept tokens, one stash per layer type
        self.kvc: Dict[str, Optional[torch.Tensor]] = {'self': None, 'cross': None}
        # high-precision attention: "lo" planes of the Q|K|V projections (and of the compacted K|V)
        self.qkv_lo: Dict[str, Optional[torch.Tensor]] = {'self': None, 'cross': None}
        self.kvc_lo: Dict[str, Optional[torch.Tensor]] = {'self': None, 'cross': None}
        # fused instance norm of the MLP hidden layer (statistics come out of the first MLP GEMM's epilogue)
        self.in_stats = ops.InstNormStats(n_img, Np, 2 * D, device) if Np >= 128 else None
        # EIMP pooling: fp32-level row LSE of the stashed attention + per-head scratch of the column sums
        self.lse_x: Dict[str, Optional[torch.Tensor]] = {'self': None, 'cross': None}
        self.cs_scratch: Optional[torch.Tensor] = None
        # keypoint encoder scratch
        self.k_in = torch.zeros(T, 4, **f32)
        self.k_a = torch.zeros(T, 64, **f32)
        self.k_b = torch.zeros(T, 64, **f32)
        self.k_c = torch.zeros(T, D, **f32)
        self.k_p = Planes.empty((T, D), device)
        self.tok_f32 = torch.zeros(T, D, **f32)

    def kv_compact(self, name: str, lo: bool = False) -> torch.Tensor:
        store = self.kvc_lo if lo else self.kvc
        if store[name] is None:
            store[name] = torch.zeros(self.n_img * self.Np, 2 * D, dtype=torch.float16, device=self.H.device)
        return store[name]

    def lse_exact(self, name: str) -> torch.Tensor:
        if self.lse_x[name] is None:
            self.lse_x[name] = torch.zeros(self.n_img, HEADS, self.Np, dtype=torch.float32, device=self.H.device)
        return self.lse_x[name]

    def colsum_scratch(self) -> torch.Tensor:
        if self.cs_scratch is None:
            self.cs_scratch = torch.empty(self.n_img, HEADS, self.Np, dtype=torch.float32, device=self.H.device)
        return self.cs_scratch

    def qkv_lo_buf(self, name: str) -> torch.Tensor:
        if self.qkv_lo[name] is None:
            self.qkv_lo[name] = torch.zeros(self.n_img * self.Np, 3 * D, dtype=torch.float16, device=self.H.device)
        return self.qkv_lo[name]


class RunState:
    """One pair batch in flight: current descriptors (planes in ws.X) plus the stashes the sharing layers need."""

    def __init__(self, ws: Workspace, B: int, N0: int, N1: int, n_tok: torch.Tensor):
        self.ws, self.B, self.N0, self.N1 = ws, B, N0, N1
        self.n_tok = n_tok              # [2B] int32 valid tokens per image
        self.key_cnt: Optional[torch.Tensor] = None   # [2B] int32 kept keys per image (EIMP), None = all
        self.key_ids: Optional[torch.Tensor] = None   # [2B, Np] int32 sorted kept ids
        self.stash_cnt = {'self': None, 'cross': None}  # key counts the stashed K (and LSE) were built with
        self.stash_ok = {'self': False, 'cross': False}   # a non-sharing layer of this type has run on THIS state
        self.ragged = False             # n_tok came from the caller (padded inputs): scoring must mask by it too


class Engine:
    def __init__(self, pk: PackedModel, names: List[str], high_precision_attention: bool = False,
                 stash_lo: bool = False):
        self.pk = pk
        self.names = names
        self.hp = high_precision_attention
        # EIMP: non-sharing layers also keep the "lo" planes of Q and K, so that the pooling statistics (attention
        # received per key) can be recomputed at fp32 level even when the layers themselves use single fp16 operands
        self.stash_lo = stash_lo and not high_precision_attention
        self._ws: Dict[tuple, Workspace] = {}

    WS_BUDGET_BYTES = 24 << 30      # cached workspaces (all shapes) stay below this; ~5.5 KB per token

    def workspace(self, n_img: int, Np: int, device) -> Workspace:
        key = (n_img, Np, str(device))
        if key not in self._ws:
            tokens = sum(k[0] * k[1] for k in self._ws) + n_img * Np
            if len(self._ws) >= 32 or tokens * 5632 > self.WS_BUDGET_BYTES:
                self._ws.clear()
            self._ws[key] = Workspace(n_img, Np, device)
        return self._ws[key]

    # ------------------------------------------------------------------ keypoint encoder
    def encode_keypoints(self, ws: Workspace, norm_kpts: torch.Tensor, scores: torch.Tensor, n_tok: torch.Tensor,
                         out_f32: torch.Tensor):
        """KeypointEncoder (nets/layers.py:80-90) on stacked images.  norm_kpts [n_img, Np, 2], scores [n_img, Np]
        -> out_f32 [n_img*Np, 256]."""
        n_img, Np = ws.n_img, ws.Np
        T = n_img * Np
        ops.kenc_input(norm_kpts, scores, ws.k_in)
        (w0, b0), (w1, b1) = self.pk.kenc_small
        ops.small_linear(ws.k_in, 4, w0, b0, ws.k_a, 64, T, 3, 32)
        ops.instnorm_relu(ws.k_a, batch=n_img, Nmax=Np, C_=64, ns=n_tok, out_f32=ws.k_b)  # 32 used columns (ld 64)
        ops.small_linear(ws.k_b, 64, w1, b1, ws.k_a, 64, T, 32, 64)
        ops.instnorm_relu(ws.k_a, batch=n_img, Nmax=Np, C_=64, ns=n_tok, out=Planes(ws.k_p.hi.view(-1)[:T * 64].view(T, 64),
                                                                                      ws.k_p.lo.view(-1)[:T * 64].view(T, 64)))
        cur_c = 64
        cur = Planes(ws.k_p.hi.view(-1)[:T * 64].view(T, 64), ws.k_p.lo.view(-1)[:T * 64].view(T, 64))
        for li, (wp, bias, shape) in enumerate(self.pk.kenc_big):
            cout = shape[0]
            last = li == len(self.pk.kenc_big) - 1
            dst = out_f32 if last else ws.k_c
            ops.gemm(cur, wp, M=T, N=cout, K1=cur_c, a_row_stride=cur_c, b_row_stride=cur_c, bias=bias,
                     out_mode=ops.OUT_F32, out0=dst, out_row_stride=cout)
            if not last:
                # normalise into a fresh plane view (never aliases the GEMM input of this step)
                nxt_hi = (ws.Hn.hi if li % 2 == 0 else ws.k_p.hi).view(-1)[:T * cout].view(T, cout)
                nxt_lo = (ws.Hn.lo if li % 2 == 0 else ws.k_p.lo).view(-1)[:T * cout].view(T, cout)
                nxt = Planes(nxt_hi, nxt_lo)
                ops.instnorm_relu(ws.k_c.view(-1)[:T * cout].view(T, cout), batch=n_img, Nmax=Np, C_=cout, ns=n_tok, out=nxt)
                cur, cur_c = nxt, cout

    # ------------------------------------------------------------------ one GNN layer on both images
    def layer(self, st: RunState, li: int):
        ws, L = st.ws, self.pk.layers[li]
        name = self.names[li]
        cross = name == 'cross'
        n_img, Np = ws.n_img, ws.Np
        T = n_img * Np
        buf = ws.qkv[name]
        lse = ws.lse[name]
        base = buf.data_ptr()
        hp = self.hp
        keep_lo = hp or (self.stash_lo and not L['sharing'])     # this call writes lo planes
        buf_lo = ws.qkv_lo_buf(name) if keep_lo else None
        base_lo = buf_lo.data_ptr() if hp else None              # ... and the attention kernel consumes them
        mode = ops.OUT_SPLIT if keep_lo else ops.OUT_F16
        if not L['sharing']:
            ops.gemm(ws.X, L['Wqkv'], M=T, N=3 * D, K1=D, a_row_stride=D, b_row_stride=D, bias=L['bqkv'],
                     out_mode=mode, out0=buf, out1=buf_lo, out_row_stride=3 * D)
            st.stash_ok[name] = True
        else:
            if not st.stash_ok[name]:
                # the reference fails with a shape error here (prob of another size / None, nets/layers.py:211-214)
                raise RuntimeError(f'sharing layer {li} ({name}) has no attention stashed for inputs of this shape: run the '
                                   f'preceding non-sharing {name} layer on the same keypoint sets first')
            ops.gemm(ws.X, L['Wv'], M=T, N=D, K1=D, a_row_stride=D, b_row_stride=D, bias=L['bv'],
                     out_mode=mode, out0=buf, out1=buf_lo, out_row_stride=3 * D, out_offset=2 * D)
        k_lo = v_lo = None
        if st.key_ids is None:
            k_ptr, v_ptr, kv_rs, nk = base + 2 * D, base + 2 * 2 * D, 3 * D, st.n_tok
            if hp:
                k_lo, v_lo = base_lo + 2 * D, base_lo + 2 * 2 * D
        else:
            planes = [(buf, ws.kv_compact(name))] + ([(buf_lo, ws.kv_compact(name, lo=True))] if keep_lo else [])
            for src, kvc in planes:
                src3 = src.view(n_img, Np, 3 * D)
                dst3 = kvc.view(n_img, Np, 2 * D)
                if not L['sharing']:
                    ops.gather_rows(src3[:, :, D:], st.key_ids, st.key_cnt, dst3, Np)          # K | V of the kept tokens
                else:
                    ops.gather_rows(src3[:, :, 2 * D:], st.key_ids, st.key_cnt, dst3[:, :, D:], Np)  # new V, same key set
            if not L['sharing']:
                st.stash_cnt[name] = st.key_cnt
            kvc = ws.kv_compact(name)
            k_ptr, v_ptr, kv_rs, nk = kvc.data_ptr(), kvc.data_ptr() + 2 * D, 2 * D, st.key_cnt
            if hp:
                kl = ws.kv_compact(name, lo=True)
                k_lo, v_lo = kl.data_ptr(), kl.data_ptr() + 2 * D
        pairs_qk = st.B * (2 * st.N0 * st.N1 if cross else st.N0 * st.N0 + st.N1 * st.N1)
        ops.ATTN_WORK_HINT = (2.0 if L['sharing'] else 4.0) * D * pairs_qk   # algorithmic FLOPs
        ops.attention(base, k_ptr, v_ptr, n_img=n_img, src_offset=(st.B if cross else 0), Nq_max=Np, Nk_max=Np,
                      nq=st.n_tok, nk=nk, shared=L['sharing'], lse=lse, out=ws.A, q_row_stride=3 * D,
                      kv_row_stride=kv_rs, q_lo=base_lo, k_lo=k_lo, v_lo=v_lo)
        ops.ATTN_WORK_HINT = None
        fused = ws.in_stats is not None and FUSED_INSTNORM
        ops.gemm(ws.X, L['W0'], M=T, N=2 * D, K1=D, K2=D, a2=ws.A, a_row_stride=D, a2_row_stride=D, b_row_stride=2 * D,
                 bias=L['b0'], out_mode=ops.OUT_F32, out0=ws.H, out_row_stride=2 * D,
                 stats=ws.in_stats if fused else None, ns=st.n_tok, Np=Np)
        if fused and FUSED_NORM_A:
            # statistics -> (mean, rstd); the normalisation itself happens in the A-operand path of the next GEMM
            ops.instnorm_apply(ws.H, ws.in_stats, batch=n_img, Nmax=Np, C_=2 * D, ns=st.n_tok, out=None)
            ops.gemm(ws.Hn, L['W3'], M=T, N=D, K1=2 * D, a_row_stride=2 * D, b_row_stride=2 * D, bias=L['b3'],
                     out_mode=ops.OUT_SPLIT_RESID, out0=ws.X.hi, out1=ws.X.lo, out_row_stride=D, res=ws.X,
                     a_f32=ws.H, a_stats=ws.in_stats, Np=Np)
            return
        if fused:
            ops.instnorm_apply(ws.H, ws.in_stats, batch=n_img, Nmax=Np, C_=2 * D, ns=st.n_tok, out=ws.Hn)
        else:
            ops.instnorm_relu(ws.H, batch=n_img, Nmax=Np, C_=2 * D, ns=st.n_tok, out=ws.Hn)
        ops.gemm(ws.Hn, L['W3'], M=T, N=D, K1=2 * D, a_row_stride=2 * D, b_row_stride=2 * D, bias=L['b3'],
                 out_mode=ops.OUT_SPLIT_RESID, out0=ws.X.hi, out1=ws.X.lo, out_row_stride=D, res=ws.X)

    def received_attention(self, st: RunState, name: str, out: torch.Tensor):
        """Un-normalised attention received by the (kept) keys of every image, indexed by key position and stored in the
        row of the KEY image (nets/adgm.py:424-427: prob.sum over heads and queries).  Recomputed from the stashed Q, K
        of the last non-sharing layer of this type.  With lo planes available (EIMP models, or attention_precision =
        'high') scores are formed at fp32 level with a matching row LSE -- the pooling rule compares these sums with
        their own median, so fp16 scores would flip tokens near it."""
        ws = st.ws
        if not st.stash_ok[name]:
            raise RuntimeError(f'no {name}-attention stashed for inputs of this shape (pool() follows forward_one_layer())')
        buf = ws.qkv[name]
        base = buf.data_ptr()
        lo = (self.hp or self.stash_lo) and ws.qkv_lo[name] is not None
        base_lo = ws.qkv_lo[name].data_ptr() if lo else None
        compact = not (st.key_ids is None or ws.kvc[name] is None or st.stash_cnt[name] is None)
        if not compact:
            k_ptr, kv_rs, nk = base + 2 * D, 3 * D, st.n_tok
            k_lo = base_lo + 2 * D if lo else None
        else:
            k_ptr, kv_rs, nk = ws.kvc[name].data_ptr(), 2 * D, st.stash_cnt[name]
            k_lo = ws.kvc_lo[name].data_ptr() if lo else None
        off = st.B if name == 'cross' else 0
        lse = ws.lse[name]
        if lo and not self.hp:
            # the stashed LSE belongs to the fp16 scores of the layer; the split-precision scores need their own
            # (LSE-only pass of the high-precision attention kernel: Q K^T with the 3-product split + softmax statistics,
            # no V, no P V)
            lse = ws.lse_exact(name)
            ops.attention(base, k_ptr, None, n_img=ws.n_img, src_offset=off, Nq_max=ws.Np, Nk_max=ws.Np, nq=st.n_tok,
                          nk=nk, shared=False, lse=lse, out=None, q_row_stride=3 * D, kv_row_stride=kv_rs, q_lo=base_lo,
                          k_lo=k_lo, v_lo=None)
        ops.attention_colsum(base, k_ptr, n_img=ws.n_img, src_offset=off, Nq_max=ws.Np, Nk_max=ws.Np, nq=st.n_tok, nk=nk,
                             lse=lse, colsum=out, q_row_stride=3 * D, kv_row_stride=kv_rs, q_lo=base_lo, k_lo=k_lo,
                             scratch=ws.colsum_scratch(), by_key_image=True)

    # ------------------------------------------------------------------ scoring
    def project(self, st: RunState, ni: int):
        """final_proj[ni] on both images (nets/gm.py:291-292) -> ws.Y planes."""
        ws = st.ws
        wf, bf = self.pk.final[ni]
        T = ws.n_img * ws.Np
        ops.gemm(ws.X, wf, M=T, N=D, K1=D, a_row_stride=D, b_row_stride=D, bias=bf, out_mode=ops.OUT_SPLIT,
                 out0=ws.Y.hi, out1=ws.Y.lo, out_row_stride=D)

    def distance(self, st: RunState, y: Planes, M: int, N: int, dist: torch.Tensor, ldd: int):
        """dist[b] = Y0[b] . Y1[b]^T / 16 (nets/gm.py:293-294); y holds [2B, Np, 256] image-major."""
        ws = st.ws
        img = ws.Np * D
        ops.gemm(y, y, M=M, N=N, K1=D, batch=st.B, a_row_stride=D, a_batch_stride=img, b_row_stride=D,
                 b_batch_stride=img, b_batched=True, b_offset=st.B * img, alpha=1.0 / 16.0, out_mode=ops.OUT_F32,
                 out0=dist, out_row_stride=ldd, out_batch_stride=dist.stride(0))
