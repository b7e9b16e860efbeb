"""ctypes binding of libimp_b200.so (C ABI declared in include/imp_b200.h).

The product path has no CPU / PyTorch fallback: if the CUDA library is missing or an entry point fails,
``ImpLibraryError`` is raised.  ``load()`` only dlopens the library (no GPU needed), so symbol checks run
on CPU-only boxes; any compute call requires a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import torch  # noqa: F401  (loads libcudart.so.12, which libimp_b200.so links against)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libimp_b200.so')
ABI_VERSION = 4


class ImpLibraryError(RuntimeError):
    pass


c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class GemmArgs(C.Structure):
    _fields_ = [('a_hi', c_vp), ('a_lo', c_vp), ('a2_hi', c_vp), ('a2_lo', c_vp),
                ('a_row_stride', c_i64), ('a_batch_stride', c_i64), ('a2_row_stride', c_i64), ('a2_batch_stride', c_i64),
                ('b_hi', c_vp), ('b_lo', c_vp), ('b_row_stride', c_i64), ('b_batch_stride', c_i64),
                ('M', c_i32), ('N', c_i32), ('K1', c_i32), ('K2', c_i32), ('batch', c_i32), ('b_batched', c_i32),
                ('nsplit', c_i32), ('alpha', c_f32), ('bias', c_vp), ('out_mode', c_i32), ('_pad', c_i32),
                ('out0', c_vp), ('out1', c_vp), ('out_row_stride', c_i64), ('out_batch_stride', c_i64),
                ('res_hi', c_vp), ('res_lo', c_vp), ('stat_partial', c_vp), ('stat_straddle', c_vp), ('stat_ns', c_vp),
                ('stat_np', c_i32), ('a_np', c_i32), ('a_f32', c_vp), ('a_stats', c_vp)]


class AttnArgs(C.Structure):
    _fields_ = [('q', c_vp), ('k', c_vp), ('v', c_vp), ('q_img_stride', c_i64), ('kv_img_stride', c_i64),
                ('q_row_stride', c_i32), ('kv_row_stride', c_i32), ('n_img', c_i32), ('src_offset', c_i32), ('Nq_max', c_i32), ('Nk_max', c_i32),
                ('nq', c_vp), ('nk', c_vp), ('shared', c_i32), ('_pad', c_i32), ('lse', c_vp),
                ('out_hi', c_vp), ('out_lo', c_vp), ('out_img_stride', c_i64),
                ('q_lo', c_vp), ('k_lo', c_vp), ('v_lo', c_vp)]


class AttnColsumArgs(C.Structure):
    _fields_ = [('q', c_vp), ('k', c_vp), ('q_img_stride', c_i64), ('kv_img_stride', c_i64),
                ('q_row_stride', c_i32), ('kv_row_stride', c_i32), ('n_img', c_i32), ('src_offset', c_i32), ('Nq_max', c_i32), ('Nk_max', c_i32),
                ('nq', c_vp), ('nk', c_vp), ('lse', c_vp), ('colsum', c_vp), ('q_lo', c_vp), ('k_lo', c_vp), ('scratch', c_vp),
                ('by_key_image', c_i32), ('_pad', c_i32)]


class SinkhornArgs(C.Structure):
    _fields_ = [('dist', c_vp), ('dist_batch_stride', c_i64), ('ldd', c_i32), ('iters', c_i32), ('bin_score', c_vp),
                ('P', c_vp), ('p_batch_stride', c_i64), ('ldp', c_i32), ('_pad', c_i32), ('u', c_vp), ('colbuf', c_vp),
                ('row_max', c_vp), ('row_arg', c_vp), ('col_key', c_vp), ('row_mass', c_vp), ('col_mass', c_vp),
                ('n0s', c_vp), ('n1s', c_vp), ('N0max', c_i32), ('N1max', c_i32), ('batch', c_i32), ('write_scores', c_i32),
                ('q_store', c_vp), ('q_batch_stride', c_i64), ('row_stats', c_vp), ('storage', c_i32), ('_pad2', c_i32)]


class MatchArgs(C.Structure):
    _fields_ = [('row_max', c_vp), ('row_arg', c_vp), ('col_key', c_vp), ('p_thresh', c_f32), ('_pad', c_i32),
                ('indices0', c_vp), ('indices1', c_vp), ('mscores0', c_vp), ('mscores1', c_vp),
                ('n0s', c_vp), ('n1s', c_vp), ('N0max', c_i32), ('N1max', c_i32), ('batch', c_i32), ('_pad2', c_i32),
                ('out0_batch_stride', c_i64), ('out1_batch_stride', c_i64)]


class PoolArgs(C.Structure):
    _fields_ = [('mass', c_vp), ('a_self', c_vp), ('a_cross', c_vp), ('ld', c_i32), ('_pad0', c_i32),
                ('ids_in', c_vp), ('cnt_in', c_vp), ('ids_out', c_vp), ('cnt_out', c_vp),
                ('changed', c_vp), ('thresh', c_f32), ('n_min_tokens', c_i32), ('batch', c_i32), ('_pad', c_i32)]


class SpConvArgs(C.Structure):
    _fields_ = [('in_hi', c_vp), ('in_lo', c_vp), ('w_hi', c_vp), ('w_lo', c_vp), ('bias', c_vp), ('out_hi', c_vp), ('out_lo', c_vp),
                ('B', c_i32), ('H', c_i32), ('W', c_i32), ('Cin', c_i32), ('Cout', c_i32), ('relu', c_i32), ('pool', c_i32), ('_pad', c_i32)]


class SpSelectArgs(C.Structure):
    _fields_ = [('scores', c_vp), ('mask', c_vp), ('H', c_i32), ('W', c_i32), ('threshold', c_f32), ('border', c_i32),
                ('max_keypoints', c_i32), ('cap', c_i32), ('rowcnt', c_vp), ('rowoff', c_vp), ('total', c_vp), ('cand_yx', c_vp),
                ('cand_score', c_vp), ('keys', c_vp), ('kpts_xy', c_vp), ('kscores', c_vp), ('n_out', c_vp)]


# name -> (restype, argtypes): every symbol include/imp_b200.h declares
SIGNATURES = {
    'imp_last_error': (C.c_char_p, []),
    'imp_abi_version': (C.c_int, []),
    'imp_set_option': (C.c_int, [c_i32, c_i32]),
    'imp_split_planes': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    'imp_merge_planes': (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    'imp_gemm': (C.c_int, [C.POINTER(GemmArgs), c_vp]),
    'imp_attention': (C.c_int, [C.POINTER(AttnArgs), c_vp]),
    'imp_attention_colsum': (C.c_int, [C.POINTER(AttnColsumArgs), c_vp]),
    'imp_instnorm_relu': (C.c_int, [c_vp, c_i64, c_i32, c_vp, c_i32, c_i32, c_i32, c_f32, c_i32, c_vp, c_vp, c_vp,
                                    c_i64, c_i32, c_vp]),
    'imp_instnorm_apply': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_f32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    'imp_kenc_input': (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    'imp_small_linear': (C.c_int, [c_vp, c_i32, c_vp, c_vp, c_vp, c_i32, c_i64, c_i32, c_i32, c_vp]),
    'imp_sinkhorn': (C.c_int, [C.POINTER(SinkhornArgs), c_vp]),
    'imp_sinkhorn_q_store_bytes': (c_i64, [c_i32, c_i32, c_i32, c_i32]),
    'imp_set_profiling': (C.c_int, [c_i32]),
    'imp_sinkhorn_iter_ms': (C.c_float, []),
    'imp_matches': (C.c_int, [C.POINTER(MatchArgs), c_vp]),
    'imp_dual_softmax': (C.c_int, [c_vp, c_i64, c_i32, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    'imp_score_argmax': (C.c_int, [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    'imp_pool_select': (C.c_int, [C.POINTER(PoolArgs), c_vp]),
    'imp_scatter_matches': (C.c_int, [c_vp, c_vp, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp]),
    'imp_gather_rows': (C.c_int, [c_vp, c_i64, c_i32, c_vp, c_i32, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp]),
    'imp_sinkhorn_rows_per_item': (C.c_int, [c_i32, c_i32, c_i32]),
    'imp_sp_conv3x3': (C.c_int, [C.POINTER(SpConvArgs), c_vp]),
    'imp_sp_conv1a': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]),
    'imp_sp_maxpool2': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    'imp_sp_scores': (C.c_int, [c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_vp]),
    'imp_sp_nms': (C.c_int, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    'imp_sp_select': (C.c_int, [C.POINTER(SpSelectArgs), c_vp]),
    'imp_sp_l2norm_rows': (C.c_int, [c_vp, c_i64, c_i32, c_vp]),
    'imp_sp_sample_descriptors': (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]),
}

_lib = None


def load():
    """dlopen libimp_b200.so and bind every declared symbol.  Raises ImpLibraryError when missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImpLibraryError(
            f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'(or `make -C imp_release_b200/csrc`).  There is no CPU fallback.')
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise ImpLibraryError(f'cannot load {LIB_PATH}: {e}') from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImpLibraryError(f'{LIB_PATH} does not export {name}') from e
        fn.restype = res
        fn.argtypes = args
    if lib.imp_abi_version() != ABI_VERSION:
        raise ImpLibraryError(f'ABI mismatch: library {lib.imp_abi_version()} != binding {ABI_VERSION}')
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().imp_last_error()
        raise ImpLibraryError(f'{what} failed (rc={rc}): {msg.decode() if msg else "?"}')


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream
