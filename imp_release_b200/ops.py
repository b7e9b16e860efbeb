"""Tensor-level wrappers around the C ABI (device memory and streams come from PyTorch; all arithmetic runs
in libimp_b200.so).  Every function enqueues on the current CUDA stream and never synchronises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (AttnArgs, AttnColsumArgs, GemmArgs, MatchArgs, PoolArgs, SinkhornArgs, check, ptr, stream_ptr)

OUT_F32, OUT_F16, OUT_SPLIT, OUT_SPLIT_RESID = 0, 1, 2, 3

# Launch accounting (bench.py): LAUNCHES counts kernels of libimp_b200.so enqueued through this module;
# when PROFILE is a dict, every wrapper brackets its launches with CUDA events on the current stream
# (name -> list of (start, end, work) tuples; `work` = algorithmic FLOPs or bytes of that call).
LAUNCHES = 0
PROFILE = None
ATTN_WORK_HINT = None   # set by callers that know the true (ragged / pruned) sizes


class _Span:
    __slots__ = ('name', 'n', 'work', 'ev')

    def __init__(self, name: str, n_kernels: int, work: float = 0.0):
        self.name, self.n, self.work, self.ev = name, n_kernels, work, None

    def __enter__(self):
        global LAUNCHES
        LAUNCHES += self.n
        if PROFILE is not None:
            self.ev = torch.cuda.Event(enable_timing=True)
            self.ev.record()
        return self

    def __exit__(self, *exc):
        if self.ev is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            PROFILE.setdefault(self.name, []).append((self.ev, end, self.work))
        return False


def on_model_device(fn):
    """Decorator for the public entry points of the host classes: run with the model's own GPU as the current CUDA device.
    The kernels launch on the CURRENT device's current stream, so a model living on cuda:1 while the caller's current device
    is cuda:0 would otherwise launch on the wrong GPU (one process driving several GPUs; per-device kernel attributes are
    handled by DeviceOnce in the launchers)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        p = next(self.parameters(), None)
        if p is None or not p.is_cuda or p.device.index == torch.cuda.current_device():
            return fn(self, *a, **kw)
        with torch.cuda.device(p.device):
            return fn(self, *a, **kw)
    return wrapped


def _require_cuda(t: torch.Tensor):
    if not t.is_cuda:
        raise _lib.ImpLibraryError('imp_release_b200 kernels need CUDA tensors (there is no CPU path)')


class Planes:
    """fp16 hi/lo plane pair representing an fp32 tensor (x ~= hi + lo, ~22 mantissa bits)."""
    __slots__ = ('hi', 'lo')

    def __init__(self, hi: torch.Tensor, lo: torch.Tensor):
        self.hi, self.lo = hi, lo

    @staticmethod
    def empty(shape, device) -> 'Planes':
        return Planes(torch.zeros(shape, dtype=torch.float16, device=device),
                      torch.zeros(shape, dtype=torch.float16, device=device))

    @property
    def shape(self):
        return self.hi.shape

    def float(self) -> torch.Tensor:
        out = torch.empty(self.hi.shape, dtype=torch.float32, device=self.hi.device)
        merge_planes(self, out)
        return out


def split_planes(x: torch.Tensor, out: Optional[Planes] = None, addend: Optional[torch.Tensor] = None) -> Planes:
    _require_cuda(x)
    x = x.contiguous()
    if addend is not None:
        addend = addend.contiguous()
    if out is None:
        out = Planes.empty(x.shape, x.device)
    with _Span('split_planes', 1, 8.0 * x.numel() * (2 if addend is not None else 1.5)):
        check(_lib.load().imp_split_planes(ptr(x), ptr(addend), ptr(out.hi), ptr(out.lo), x.numel(), stream_ptr()),
              'imp_split_planes')
    return out


def merge_planes(p: Planes, out: torch.Tensor) -> torch.Tensor:
    with _Span('merge_planes', 1):
        check(_lib.load().imp_merge_planes(ptr(p.hi), ptr(p.lo), ptr(out), p.hi.numel(), stream_ptr()), 'imp_merge_planes')
    return out


def gemm(a: Planes, b: Planes, *, M: int, N: int, K1: int, batch: int = 1, a_row_stride: int, a_batch_stride: int = 0,
         b_row_stride: int, b_batch_stride: int = 0, b_batched: bool = False, a2: Optional[Planes] = None, K2: int = 0,
         a2_row_stride: int = 0, a2_batch_stride: int = 0, nsplit: int = 3, alpha: float = 1.0,
         bias: Optional[torch.Tensor] = None, out_mode: int = OUT_F32, out0: torch.Tensor = None,
         out1: Optional[torch.Tensor] = None, out_row_stride: int = 0, out_batch_stride: int = 0,
         res: Optional[Planes] = None, a_offset: int = 0, a2_offset: int = 0, b_offset: int = 0, out_offset: int = 0,
         stats: Optional['InstNormStats'] = None, ns: Optional[torch.Tensor] = None, Np: int = 0,
         a_f32: Optional[torch.Tensor] = None, a_stats: Optional['InstNormStats'] = None):
    """D = alpha * A.B^T (+bias)(+res).  Offsets are in elements from the start of the plane tensors.
    stats / ns / Np: fused instance-norm statistics of the fp32 output (see imp_gemm_args).
    a_f32 / a_stats (+ Np): A = relu(instance_norm(a_f32)) formed inside the kernel (`a` is ignored)."""
    g = GemmArgs()
    esz = 2
    if a_f32 is not None:
        g.a_f32, g.a_stats, g.a_np = ptr(a_f32), ptr(a_stats.stats), Np
    else:
        g.a_hi = a.hi.data_ptr() + a_offset * esz
        g.a_lo = a.lo.data_ptr() + a_offset * esz
    if a2 is not None:
        g.a2_hi = a2.hi.data_ptr() + a2_offset * esz
        g.a2_lo = a2.lo.data_ptr() + a2_offset * esz
    g.a_row_stride, g.a_batch_stride = a_row_stride, a_batch_stride
    g.a2_row_stride, g.a2_batch_stride = a2_row_stride, a2_batch_stride
    g.b_hi = b.hi.data_ptr() + b_offset * esz
    g.b_lo = b.lo.data_ptr() + b_offset * esz
    g.b_row_stride, g.b_batch_stride = b_row_stride, b_batch_stride
    g.M, g.N, g.K1, g.K2, g.batch, g.b_batched = M, N, K1, K2, batch, int(b_batched)
    g.nsplit, g.alpha = nsplit, alpha
    g.bias = ptr(bias)
    g.out_mode = out_mode
    osz = 4 if out_mode == OUT_F32 else 2
    g.out0 = out0.data_ptr() + out_offset * osz
    g.out1 = (out1.data_ptr() + out_offset * osz) if out1 is not None else None
    g.out_row_stride, g.out_batch_stride = out_row_stride, out_batch_stride
    if res is not None:
        g.res_hi = res.hi.data_ptr() + out_offset * esz
        g.res_lo = res.lo.data_ptr() + out_offset * esz
    if stats is not None:
        g.stat_partial, g.stat_straddle, g.stat_ns, g.stat_np = ptr(stats.partial), ptr(stats.straddle), ptr(ns), Np
    with _Span(f'gemm_n{N}_k{K1 + K2}' + ('_b' if b_batched else ''), 1, 2.0 * M * N * (K1 + K2) * batch):
        check(_lib.load().imp_gemm(C.byref(g), stream_ptr()), 'imp_gemm')


def _p(t):
    """tensor or raw integer device address"""
    return t if isinstance(t, int) or t is None else t.data_ptr()


def attention(q, k, v, *, n_img: int, src_offset: int, Nq_max: int, Nk_max: int, nq, nk, shared: bool, lse,
              out: Optional[Planes], q_row_stride: int = 256, kv_row_stride: int = 256, q_img_stride: Optional[int] = None,
              kv_img_stride: Optional[int] = None, q_lo=None, k_lo=None, v_lo=None):
    """q/k/v: fp16 tensors or raw device addresses (slices of a fused projection buffer).
    ``work_hint``: algorithmic FLOPs of this call (QK^T + PV = 4*d per score element), for the profiler only."""
    attn_work = ATTN_WORK_HINT if ATTN_WORK_HINT is not None else 4.0 * 64 * 4 * n_img * Nq_max * Nk_max
    a = AttnArgs()
    a.q, a.k, a.v = _p(q), _p(k), _p(v)
    a.q_row_stride, a.kv_row_stride = q_row_stride, kv_row_stride
    a.q_img_stride = q_img_stride if q_img_stride is not None else Nq_max * q_row_stride
    a.kv_img_stride = kv_img_stride if kv_img_stride is not None else Nk_max * kv_row_stride
    a.n_img, a.src_offset, a.Nq_max, a.Nk_max = n_img, src_offset, Nq_max, Nk_max
    a.nq, a.nk = ptr(nq), ptr(nk)
    a.shared = int(shared)
    a.lse = ptr(lse)
    a.out_hi, a.out_lo = (ptr(out.hi), ptr(out.lo)) if out is not None else (None, None)   # None: LSE-only pass
    a.out_img_stride = Nq_max * 256
    a.q_lo, a.k_lo, a.v_lo = _p(q_lo), _p(k_lo), _p(v_lo)
    with _Span('attention_lse_only' if out is None else 'attention_shared' if shared else 'attention', 1, attn_work):
        check(_lib.load().imp_attention(C.byref(a), stream_ptr()), 'imp_attention')


def attention_colsum(q, k, *, n_img: int, src_offset: int, Nq_max: int, Nk_max: int, nq, nk, lse, colsum,
                     q_row_stride: int = 256, kv_row_stride: int = 256, q_img_stride: Optional[int] = None,
                     kv_img_stride: Optional[int] = None, q_lo=None, k_lo=None, scratch: Optional[torch.Tensor] = None,
                     by_key_image: bool = False):
    """Attention received per key.  Row `img` of `colsum` belongs to the QUERY image img (its keys are those of image
    (img + src_offset) % n_img) unless by_key_image.  q_lo / k_lo: split-precision scores (fp32 level)."""
    a = AttnColsumArgs()
    a.q, a.k = _p(q), _p(k)
    a.q_row_stride, a.kv_row_stride = q_row_stride, kv_row_stride
    a.q_img_stride = q_img_stride if q_img_stride is not None else Nq_max * q_row_stride
    a.kv_img_stride = kv_img_stride if kv_img_stride is not None else Nk_max * kv_row_stride
    a.n_img, a.src_offset, a.Nq_max, a.Nk_max = n_img, src_offset, Nq_max, Nk_max
    a.nq, a.nk = ptr(nq), ptr(nk)
    a.lse, a.colsum = ptr(lse), ptr(colsum)
    a.q_lo, a.k_lo = _p(q_lo), _p(k_lo)
    if scratch is None:
        scratch = torch.empty(n_img, 4, Nk_max, dtype=torch.float32, device=colsum.device)
    a.scratch = ptr(scratch)
    a.by_key_image = int(by_key_image)
    with _Span('attention_colsum', 2, (3 if q_lo is not None else 1) * 2.0 * 64 * 4 * n_img * Nq_max * Nk_max):
        check(_lib.load().imp_attention_colsum(C.byref(a), stream_ptr()), 'imp_attention_colsum')


def instnorm_relu(H: torch.Tensor, *, batch: int, Nmax: int, C_: int, ns=None, eps: float = 1e-3, relu: bool = True,
                  out: Optional[Planes] = None, out_f32: Optional[torch.Tensor] = None):
    """H [batch, Nmax, C] fp32 contiguous -> planes / fp32 of the same logical shape."""
    with _Span(f'instnorm_c{C_}', 1, 8.0 * batch * Nmax * C_):
        check(_lib.load().imp_instnorm_relu(ptr(H), Nmax * C_, C_, ptr(ns), Nmax, C_, batch, eps, int(relu),
                                            ptr(out.hi) if out is not None else None,
                                            ptr(out.lo) if out is not None else None, ptr(out_f32), Nmax * C_, C_,
                                            stream_ptr()), 'imp_instnorm_relu')


class InstNormStats:
    """Workspace of the fused instance norm: per-tile partial sums written by the GEMM epilogue, reduced by
    instnorm_apply into (mean, rstd) per (image, channel)."""

    def __init__(self, n_img: int, Np: int, C_: int, device):
        tiles = (n_img * Np + 127) // 128
        f32 = dict(dtype=torch.float32, device=device)
        self.partial = torch.zeros(tiles, C_, 2, **f32)
        self.straddle = torch.zeros(n_img, 4, C_, 2, **f32)
        self.stats = torch.zeros(n_img, C_, 2, **f32)


def instnorm_apply(H: torch.Tensor, st: InstNormStats, *, batch: int, Nmax: int, C_: int, ns, out: Optional[Planes],
                   eps: float = 1e-3, relu: bool = True):
    """Second half of the fused instance norm: statistics left by gemm(..., stats=st) -> (mean, rstd) in st.stats, then
    normalise + ReLU + hi/lo split into `out` -- or, with out=None, nothing more (the consumer GEMM normalises its A operand
    itself: gemm(..., a_f32=H, a_stats=st))."""
    name = f'instnorm_apply_c{C_}' if out is not None else 'instnorm_finalize'
    with _Span(name, 2 if out is not None else 1, 8.0 * batch * Nmax * C_ if out is not None else 0.0):
        check(_lib.load().imp_instnorm_apply(ptr(H), ptr(st.partial), ptr(st.straddle), ptr(ns), Nmax, C_, batch, eps, int(relu),
                                             ptr(st.stats), ptr(out.hi) if out is not None else None,
                                             ptr(out.lo) if out is not None else None, stream_ptr()), 'imp_instnorm_apply')


def kenc_input(norm_kpts: torch.Tensor, scores: torch.Tensor, out: torch.Tensor):
    with _Span('kenc_input', 1):
        check(_lib.load().imp_kenc_input(ptr(norm_kpts), ptr(scores), ptr(out), scores.numel(), stream_ptr()),
              'imp_kenc_input')


def small_linear(X, ldx, W, bias, Y, ldy, rows, cin, cout):
    with _Span('small_linear', 1, 2.0 * rows * cin * cout):
        check(_lib.load().imp_small_linear(ptr(X), ldx, ptr(W), ptr(bias), ptr(Y), ldy, rows, cin, cout, stream_ptr()),
              'imp_small_linear')


SK_STORAGE = {'fp32': 0, 'fp16': 1, 'fp24': 2}
SK_NO_RESIDENT = 0x100


def default_sk_storage() -> str:
    """Storage of softmax(M) for the Sinkhorn iteration sweeps (include/imp_b200.h, IMP_SK_STORE_*).  'fp32' is the
    reference recurrence on exact fp32 probabilities (match indices bit-exact with the reference on every fixture);
    'fp24' / 'fp16' move fewer bytes per sweep at the price of ~1e-5 / ~1e-3 relative noise on the scaling vectors
    (opt-in: a near-tie below that level can pick the other candidate)."""
    return os.environ.get('IMP_SK_STORAGE', 'fp32')


OPT_SK_RESIDENT, OPT_ATTN_VARIANT, OPT_GEMM_VARIANT, OPT_SM_LIMIT = 1, 2, 3, 4


def set_option(key: int, value: int):
    check(_lib.load().imp_set_option(key, value), 'imp_set_option')


def set_sinkhorn_resident(on: bool):
    """False: small Sinkhorn problems take the streaming kernels of the big batches too (parity tests of that path).
    Workspaces built before the switch keep their old geometry -- create them afterwards."""
    set_option(OPT_SK_RESIDENT, int(bool(on)))


class SinkhornWorkspace:
    """Buffers for one Sinkhorn + matching call on [batch, N0max, N1max] problems."""

    def __init__(self, batch: int, N0max: int, N1max: int, device, want_mass: bool = False, storage: Optional[str] = None,
                 resident: bool = True):
        self.batch, self.N0max, self.N1max = batch, N0max, N1max
        self.ldp = (N1max + 1 + 3) // 4 * 4
        self.storage = SK_STORAGE[storage if storage is not None else default_sk_storage()]
        if not resident:
            self.storage |= SK_NO_RESIDENT      # streaming kernels only (no cooperative launch)
        f32 = dict(dtype=torch.float32, device=device)
        self.P = torch.zeros(batch, N0max + 1, self.ldp, **f32)
        self.u = torch.empty(batch, N0max + 1, **f32)
        self.colbuf = torch.empty(3, batch, self.ldp, **f32)
        self.row_max = torch.zeros(batch, N0max, **f32)
        self.row_arg = torch.zeros(batch, N0max, dtype=torch.int32, device=device)
        self.col_key = torch.zeros(batch, N1max, dtype=torch.int64, device=device)
        self.row_mass = torch.zeros(batch, N0max, **f32) if want_mass else None
        self.col_mass = torch.zeros(batch, N1max, **f32) if want_mass else None
        self.q_store = self.row_stats = None
        self.q_batch_stride = 0
        legacy = os.environ.get('IMP_SK_LEGACY', '0') == '1'      # row-ring fp32 kernels of csrc/sinkhorn.cu
        per_matrix = 0 if legacy else int(_lib.load().imp_sinkhorn_q_store_bytes(batch, N0max, N1max, self.storage))
        if per_matrix > 0:
            self.q_batch_stride = per_matrix
            self.q_store = torch.empty(batch, per_matrix, dtype=torch.uint8, device=device)
            self.row_stats = torch.empty(2, batch, N0max + 1, **f32)

    def scores(self) -> torch.Tensor:
        """[batch, N0max+1, N1max+1] view of the padded score buffer (a real torch.Tensor)."""
        return self.P[:, :, :self.N1max + 1]


def sinkhorn(dist: torch.Tensor, ldd: int, bin_score: torch.Tensor, iters: int, ws: SinkhornWorkspace, n0s=None,
             n1s=None, dist_batch_stride: Optional[int] = None, write_scores: bool = True):
    a = SinkhornArgs()
    a.dist = ptr(dist)
    a.dist_batch_stride = dist_batch_stride if dist_batch_stride is not None else ws.N0max * ldd
    a.ldd, a.iters = ldd, iters
    a.bin_score = ptr(bin_score)
    a.P, a.p_batch_stride, a.ldp = ptr(ws.P), (ws.N0max + 1) * ws.ldp, ws.ldp
    a.u, a.colbuf = ptr(ws.u), ptr(ws.colbuf)
    a.row_max, a.row_arg, a.col_key = ptr(ws.row_max), ptr(ws.row_arg), ptr(ws.col_key)
    a.row_mass, a.col_mass = ptr(ws.row_mass), ptr(ws.col_mass)
    a.n0s, a.n1s = ptr(n0s), ptr(n1s)
    a.N0max, a.N1max, a.batch = ws.N0max, ws.N1max, ws.batch
    a.write_scores = int(write_scores)
    a.q_store, a.q_batch_stride, a.row_stats, a.storage = ptr(ws.q_store), ws.q_batch_stride, ptr(ws.row_stats), ws.storage
    mat_bytes = 4.0 * ws.batch * (ws.N0max + 1) * (ws.N1max + 1)
    # algorithmic traffic: 2 sweeps per iteration + init (read dist, write p) + final (read, write)
    n_kernels = (2 if ws.q_store is not None else 3) + max(iters - 1, 0)     # init, iterations, final (+ column arg-max)
    with _Span('sinkhorn', n_kernels, mat_bytes * (2 * iters + 4)):
        check(_lib.load().imp_sinkhorn(C.byref(a), stream_ptr()), 'imp_sinkhorn')


def matches(ws_row_max, ws_row_arg, ws_col_key, p: float, N0max: int, N1max: int, batch: int, n0s=None, n1s=None,
            want1: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor, Optional[torch.Tensor]]:
    dev = ws_row_max.device
    # the kernel writes every entry (-1 / 0 beyond a sample's size): no fill launches
    i0 = torch.empty((batch, N0max), dtype=torch.int64, device=dev)
    m0 = torch.empty(batch, N0max, dtype=torch.float32, device=dev)
    i1 = torch.empty((batch, N1max), dtype=torch.int64, device=dev) if want1 else None
    m1 = torch.empty(batch, N1max, dtype=torch.float32, device=dev) if want1 else None
    m = MatchArgs()
    m.row_max, m.row_arg, m.col_key = ptr(ws_row_max), ptr(ws_row_arg), ptr(ws_col_key)
    m.p_thresh = p
    m.indices0, m.indices1, m.mscores0, m.mscores1 = ptr(i0), ptr(i1), ptr(m0), ptr(m1)
    m.n0s, m.n1s = ptr(n0s), ptr(n1s)
    m.N0max, m.N1max, m.batch = N0max, N1max, batch
    m.out0_batch_stride, m.out1_batch_stride = N0max, N1max
    with _Span('matches', 1):
        check(_lib.load().imp_matches(C.byref(m), stream_ptr()), 'imp_matches')
    return i0, i1, m0, m1


def score_argmax(P: torch.Tensor, N0: int, N1: int, want_mass: bool = False, n0s=None, n1s=None):
    """Row/col arg-max (and masses) over P[:, :N0, :N1] (per sample: [:n0s[b], :n1s[b]]) for an arbitrary (possibly
    strided-row) fp32 score tensor."""
    batch = P.shape[0]
    if P.stride(2) != 1:
        P = P.contiguous()
    dev = P.device
    mk = torch.zeros if n0s is not None else torch.empty        # rows / columns beyond a sample's size stay 0
    row_max = mk(batch, N0, dtype=torch.float32, device=dev)
    row_arg = mk(batch, N0, dtype=torch.int32, device=dev)
    col_key = torch.empty(batch, N1, dtype=torch.int64, device=dev)
    row_mass = mk(batch, N0, dtype=torch.float32, device=dev) if want_mass else None
    col_mass = torch.empty(batch, N1, dtype=torch.float32, device=dev) if want_mass else None
    with _Span('score_argmax', 2):
        check(_lib.load().imp_score_argmax(ptr(P), P.stride(0), P.stride(1), ptr(row_max), ptr(row_arg), ptr(col_key),
                                           ptr(row_mass), ptr(col_mass), N0, N1, batch, ptr(n0s), ptr(n1s), stream_ptr()),
              'imp_score_argmax')
    if want_mass:
        return row_max, row_arg, col_key, row_mass, col_mass
    return row_max, row_arg, col_key


def dual_softmax(dist: torch.Tensor, ldd: int, bin_score: torch.Tensor, N0: int, N1: int, batch: int, n0s=None, n1s=None,
                 dist_batch_stride: Optional[int] = None):
    dev = dist.device
    ldp = (N1 + 1 + 3) // 4 * 4
    P = torch.empty(batch, N0 + 1, ldp, dtype=torch.float32, device=dev)
    row_lse = torch.empty(batch, N0 + 1, dtype=torch.float32, device=dev)
    col_lse = torch.empty(batch, N1 + 1, dtype=torch.float32, device=dev)
    d_bs = dist_batch_stride if dist_batch_stride is not None else N0 * ldd
    with _Span('dual_softmax', 3):
        check(_lib.load().imp_dual_softmax(ptr(dist), d_bs, ldd, ptr(bin_score), ptr(P), (N0 + 1) * ldp, ldp,
                                           ptr(row_lse), ptr(col_lse), N0, N1, batch, ptr(n0s), ptr(n1s), stream_ptr()),
              'imp_dual_softmax')
    return P[:, :, :N1 + 1]


def pool_select(mass, a_self, a_cross, ids_in, cnt_in, thresh: float, n_min_tokens: int):
    """All of mass / a_self / a_cross / ids_in are [batch, ld] and indexed by subset position."""
    batch, ld = ids_in.shape
    dev = ids_in.device
    assert mass.stride(0) == ld and a_self.stride(0) == ld and a_cross.stride(0) == ld
    ids_out = torch.zeros_like(ids_in)
    cnt_out = torch.zeros(batch, dtype=torch.int32, device=dev)
    changed = torch.zeros(batch, dtype=torch.int32, device=dev)
    a = PoolArgs()
    a.mass, a.a_self, a.a_cross = ptr(mass), ptr(a_self), ptr(a_cross)
    a.ld = ld
    a.ids_in, a.cnt_in = ptr(ids_in), ptr(cnt_in)
    a.ids_out, a.cnt_out, a.changed = ptr(ids_out), ptr(cnt_out), ptr(changed)
    a.thresh, a.n_min_tokens, a.batch = thresh, n_min_tokens, batch
    with _Span('pool_select', 1):
        check(_lib.load().imp_pool_select(C.byref(a), stream_ptr()), 'imp_pool_select')
    return ids_out, cnt_out, changed


def scatter_matches(idx0, ms0, gids0, gids1, cnt0, out_idx, out_ms):
    batch, ld_sub = idx0.shape
    with _Span('scatter_matches', 1):
        check(_lib.load().imp_scatter_matches(ptr(idx0), ptr(ms0), ld_sub, ptr(gids0), ptr(gids1), gids0.stride(0), ptr(cnt0),
                                              ptr(out_idx), ptr(out_ms), out_idx.stride(0), batch, stream_ptr()),
              'imp_scatter_matches')


def gather_rows(src: torch.Tensor, ids: torch.Tensor, cnt: torch.Tensor, out: torch.Tensor, max_rows: int):
    """out[b, r, :] = src[b, ids[b, r], :] for r < cnt[b]; src/out [batch, rows, C] contiguous, same dtype."""
    batch = src.shape[0]
    rb_in = src.stride(1) * src.element_size()
    rb_out = out.stride(1) * out.element_size()
    copy = src.shape[2] * src.element_size()
    with _Span('gather_rows', 1):
        check(_lib.load().imp_gather_rows(ptr(src), src.stride(0) * src.element_size(), rb_in, ptr(ids), ids.stride(0),
                                          ptr(cnt), ptr(out), out.stride(0) * out.element_size(), rb_out, copy, max_rows,
                                          batch, stream_ptr()), 'imp_gather_rows')


# ---------------------------------------------------------------------------------------------------------------
# SuperPoint front-end (csrc/superpoint.cu; reference nets/superpoint.py)
def sp_conv3x3(x: Planes, w: Planes, bias: torch.Tensor, out: Planes, relu: bool = True, pool: bool = False):
    """x planes [B, H, W, Cin] -> out planes [B, H, W, Cout] ([B, H/2, W/2, Cout] with the fused 2 x 2 max pooling);
    w planes [Cout, 9 * Cin] (tap-major)."""
    B, H, W, Cin = x.hi.shape
    Cout = w.hi.shape[0]
    a = _lib.SpConvArgs()
    a.in_hi, a.in_lo, a.w_hi, a.w_lo, a.bias = ptr(x.hi), ptr(x.lo), ptr(w.hi), ptr(w.lo), ptr(bias)
    a.out_hi, a.out_lo = ptr(out.hi), ptr(out.lo)
    a.B, a.H, a.W, a.Cin, a.Cout, a.relu, a.pool = B, H, W, Cin, Cout, int(relu), int(pool)
    with _Span(f'sp_conv3x3_{Cin}_{Cout}', 1, 2.0 * B * H * W * 9 * Cin * Cout):
        check(_lib.load().imp_sp_conv3x3(C.byref(a), stream_ptr()), 'imp_sp_conv3x3')
    return out


def sp_conv1a(img: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, out: Planes):
    B, H, W = img.shape
    with _Span('sp_conv1a', 1):
        check(_lib.load().imp_sp_conv1a(ptr(img), ptr(w), ptr(bias), ptr(out.hi), ptr(out.lo), B, H, W, stream_ptr()), 'imp_sp_conv1a')
    return out


def sp_maxpool2(x: Planes, out: Planes):
    B, H, W, Cc = x.hi.shape
    with _Span('sp_maxpool2', 1):
        check(_lib.load().imp_sp_maxpool2(ptr(x.hi), ptr(x.lo), ptr(out.hi), ptr(out.lo), B, H, W, Cc, stream_ptr()), 'imp_sp_maxpool2')
    return out


def sp_scores(logits: torch.Tensor, scores: torch.Tensor, B: int, Hc: int, Wc: int):
    with _Span('sp_scores', 1):
        check(_lib.load().imp_sp_scores(ptr(logits), logits.shape[-1], ptr(scores), B, Hc, Wc, stream_ptr()), 'imp_sp_scores')
    return scores


def sp_nms(scores: torch.Tensor, mask: torch.Tensor, supp: torch.Tensor, radius: int):
    B, H, W = scores.shape
    with _Span('sp_nms', 5):
        check(_lib.load().imp_sp_nms(ptr(scores), ptr(mask), ptr(supp), B, H, W, radius, stream_ptr()), 'imp_sp_nms')
    return mask


class SpSelectWorkspace:
    """Scratch + outputs of the keypoint selection of one H x W image."""

    def __init__(self, H: int, W: int, max_keypoints: int, device):
        self.H, self.W = H, W
        self.cap = H * W
        np2 = 1
        while np2 < self.cap:
            np2 <<= 1
        i32 = dict(dtype=torch.int32, device=device)
        self.rowcnt, self.rowoff = torch.zeros(H, **i32), torch.zeros(H, **i32)
        self.total, self.n_out = torch.zeros(1, **i32), torch.zeros(1, **i32)
        self.cand_yx = torch.zeros(self.cap, 2, **i32)
        self.cand_score = torch.zeros(self.cap, dtype=torch.float32, device=device)
        self.keys = torch.zeros(np2, dtype=torch.int64, device=device)
        self.max_out = self.cap if max_keypoints < 0 else min(self.cap, max_keypoints)
        self.kpts = torch.zeros(max(self.max_out, 1), 2, dtype=torch.float32, device=device)
        self.kscores = torch.zeros(max(self.max_out, 1), dtype=torch.float32, device=device)


def sp_select(scores: torch.Tensor, mask: torch.Tensor, ws: SpSelectWorkspace, threshold: float, border: int, max_keypoints: int):
    """scores / mask: [H, W] of one image.  Results in ws.kpts[:n], ws.kscores[:n] with n = ws.n_out (device)."""
    a = _lib.SpSelectArgs()
    a.scores, a.mask, a.H, a.W = ptr(scores), ptr(mask), ws.H, ws.W
    a.threshold, a.border, a.max_keypoints, a.cap = float(threshold), int(border), int(max_keypoints), ws.cap
    a.rowcnt, a.rowoff, a.total = ptr(ws.rowcnt), ptr(ws.rowoff), ptr(ws.total)
    a.cand_yx, a.cand_score, a.keys = ptr(ws.cand_yx), ptr(ws.cand_score), ptr(ws.keys)
    a.kpts_xy, a.kscores, a.n_out = ptr(ws.kpts), ptr(ws.kscores), ptr(ws.n_out)
    with _Span('sp_select', 4):
        check(_lib.load().imp_sp_select(C.byref(a), stream_ptr()), 'imp_sp_select')


def sp_l2norm_rows(x: torch.Tensor):
    rows = x.numel() // x.shape[-1]
    with _Span('sp_l2norm_rows', 1):
        check(_lib.load().imp_sp_l2norm_rows(ptr(x), rows, x.shape[-1], stream_ptr()), 'imp_sp_l2norm_rows')
    return x


def sp_sample_descriptors(dmap: torch.Tensor, kpts: torch.Tensor, n_kpts: Optional[torch.Tensor], out: torch.Tensor, Hc: int, Wc: int,
                          max_k: int):
    with _Span('sp_sample_descriptors', 1):
        check(_lib.load().imp_sp_sample_descriptors(ptr(dmap), ptr(kpts), ptr(n_kpts), ptr(out), Hc, Wc, max_k, stream_ptr()),
              'imp_sp_sample_descriptors')
    return out
