/* imp_b200.h -- C ABI of libimp_b200.so: the sm_100a kernels behind the IMP / EIMP matcher hot path.
 *
 * Drop-in boundary: the reference (feixue94/imp-release) is pure PyTorch; its hot path is the set of
 * nn.Module methods in nets/layers.py, nets/gm.py, nets/gms.py, nets/adgm.py that eval/eval_imp.py and
 * eval/matching.py call.  The host side of this repo (imp_release_b200/nets/*.py, re-exported as `nets.*` by
 * dropin/) keeps those Python signatures and state_dict keys and binds the entry points below through ctypes
 * (INTEGRATION.md shows the binding).  Every entry point cites the reference lines it replaces.
 *
 * Conventions: all pointers are DEVICE pointers unless noted; the caller (PyTorch) owns every buffer; nothing is
 * allocated inside; every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns 0 on success,
 * non-zero on error with a message available from imp_last_error() (thread local).  Token tensors are token-major
 * [images, N, C].  "hi/lo planes" = two fp16 tensors whose sum carries ~22 mantissa bits of an fp32 value; the
 * tensor-core GEMMs consume them with a 3-product split (hi*hi + lo*hi + hi*lo) for fp32-level accuracy.
 */
#ifndef IMP_B200_H_
#define IMP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMP_B200_ABI_VERSION 4
#if defined(__GNUC__)
#define IMP_API __attribute__((visibility("default")))
#else
#define IMP_API
#endif

IMP_API const char* imp_last_error(void);
IMP_API int imp_abi_version(void);

/* Run-time knobs of the library (process-wide; no reference counterpart -- the reference has no kernels to tune).
 * IMP_OPT_SK_RESIDENT: 1 (default) = small Sinkhorn problems run the shared-memory-resident kernel, 0 = every problem
 * takes the streaming kernels the big batches use (what the parity tests switch on to cover that path with the
 * reference fixtures).  IMP_OPT_ATTN_VARIANT / IMP_OPT_GEMM_VARIANT: kernel variants (0 = default), for tuning runs.
 * IMP_OPT_SM_LIMIT: number of SMs the following launches may assume (0 = the whole device).  The host layer sets it before it
 * launches onto a stream that belongs to an SM partition (CUDA green context), so that persistent kernels size their grids for
 * the partition instead of the device. */
enum { IMP_OPT_SK_RESIDENT = 1, IMP_OPT_ATTN_VARIANT = 2, IMP_OPT_GEMM_VARIANT = 3, IMP_OPT_SM_LIMIT = 4 };
IMP_API int imp_set_option(int32_t key, int32_t value);

/* ---- fp32 <-> hi/lo planes (boundary conversions; `addend` may be NULL) -------------------------------------- */
/* desc + keypoint encoding, nets/gms.py:171-172 */
IMP_API int imp_split_planes(const float* x, const float* addend, void* hi, void* lo, int64_t n, void* stream);
IMP_API int imp_merge_planes(const void* hi, const void* lo, float* x, int64_t n, void* stream);

/* ---- split-precision tensor-core GEMM: D = alpha * A . B^T (+bias) (+residual) -------------------------------
 * replaces every Conv1d(kernel_size=1): nets/layers.py:119 (Q/K/V proj), :134 (merge, folded into W0),
 * :149/:218 + :59-77 (MLP convs), nets/gm.py:291-292 (final_proj) and the score einsum nets/gm.py:293-294. */
enum { IMP_GEMM_OUT_F32 = 0, IMP_GEMM_OUT_F16 = 1, IMP_GEMM_OUT_SPLIT = 2, IMP_GEMM_OUT_SPLIT_RESID = 3 };

typedef struct imp_gemm_args {
  const void *a_hi, *a_lo, *a2_hi, *a2_lo; /* A [batch, M, K1] (+ second K segment [batch, M, K2]) fp16 planes */
  int64_t a_row_stride, a_batch_stride, a2_row_stride, a2_batch_stride; /* in elements */
  const void *b_hi, *b_lo; /* B [b_batched ? batch : 1, N, K1+K2] fp16 planes */
  int64_t b_row_stride, b_batch_stride;
  int32_t M, N, K1, K2, batch, b_batched;
  int32_t nsplit; /* 3 = hi/lo split (fp32-level accuracy), 1 = hi planes only */
  float alpha;
  const float* bias; /* [N] or NULL */
  int32_t out_mode;
  int32_t _pad;
  void *out0, *out1; /* F32: out0 fp32 | F16: out0 fp16 | SPLIT*: out0 = hi plane, out1 = lo plane */
  int64_t out_row_stride, out_batch_stride;
  const void *res_hi, *res_lo; /* SPLIT_RESID: residual planes, same strides as out (may alias out) */
  /* Optional fused InstanceNorm1d statistics of the output (nets/layers.py:68-72; OUT_F32, N > 128, batch = 1): the M rows
   * are images of stat_np >= 128 rows each, of which the first stat_ns[img] are real tokens.  The epilogue leaves, per
   * 128-row tile and column, (sum, sum of squares) over the tile's valid rows of its FIRST image in stat_partial
   * [ceil(M/128)][N][2], and for tiles reaching into the next image the four 32-row slab sums over that image's rows in
   * stat_straddle [images][4][N][2].  Plain stores in fixed order: imp_instnorm_apply reduces them deterministically. */
  float* stat_partial;
  float* stat_straddle;
  const int32_t* stat_ns;
  int32_t stat_np;
  /* Optional normalising A operand (the consumer side of the fused instance norm, nets/layers.py:68-72 -> :73): with a_f32
   * != NULL the A planes are ignored and A = relu((a_f32 - mean) * rstd), formed in the kernel from the fp32 matrix
   * a_f32 [M, K1] (row stride a_row_stride floats) and a_stats [M / a_np][K1][2] = (mean, rstd) per image and channel
   * (what imp_instnorm_apply's finalize step leaves in `stats`).  One K segment, nsplit = 3, N > 128, batch = 1. */
  int32_t a_np;
  const float* a_f32;
  const float* a_stats;
} imp_gemm_args;
IMP_API int imp_gemm(const imp_gemm_args* args, void* stream);

/* ---- multi-head attention core (QK^T softmax PV), nets/layers.py:121-131 and :211-214 ----------------------- */
typedef struct imp_attn_args {
  const void *q, *k, *v; /* fp16 [n_img, N*_max, (row stride)], 256 used columns, heads contiguous */
  int64_t q_img_stride, kv_img_stride; /* elements */
  int32_t q_row_stride, kv_row_stride; /* elements (256 for packed tensors, 768 inside a fused QKV buffer) */
  int32_t n_img, src_offset, Nq_max, Nk_max;
  const int32_t *nq, *nk; /* [n_img] valid rows (NULL -> max) */
  int32_t shared;         /* 1: reuse previous probabilities via lse (SharedAttentionalPropagation) */
  int32_t _pad;
  float* lse;            /* [n_img, 4, Nq_max] log2-domain LSE; written if !shared, read if shared */
  void *out_hi, *out_lo; /* [n_img, Nq_max, 256] fp16 planes.  Both NULL (with q_lo, k_lo given, shared = 0): LSE-only pass --
                          * just the split-precision row LSE is produced (no V, no P V; v / v_lo may be NULL): what EIMP's pooling
                          * statistics need next to imp_attention_colsum (nets/adgm.py:424-427) */
  int64_t out_img_stride;
  /* optional "lo" planes of Q, K, V (same strides).  All three non-NULL selects the high-precision mode: 3-product
   * split contractions for QK^T and PV (fp32-level attention, ~1.6x the time); NULL = single fp16 operands */
  const void *q_lo, *k_lo, *v_lo;
} imp_attn_args;
IMP_API int imp_attention(const imp_attn_args* args, void* stream);

/* attention received per source token, nets/adgm.py:424-427, :557-560 (prob.sum(1).sum(1)).  Deterministic (no
 * floating-point atomics): per-head sums go to `scratch`, a second kernel adds the four heads in fixed order.
 * With q_lo / k_lo (fp16 "lo" planes of Q and K, same strides) the scores are formed with the 3-product split, i.e. at
 * fp32 level -- pass an `lse` computed the same way (imp_attention in high-precision mode writes one). */
typedef struct imp_attn_colsum_args {
  const void *q, *k;
  int64_t q_img_stride, kv_img_stride;
  int32_t q_row_stride, kv_row_stride;
  int32_t n_img, src_offset, Nq_max, Nk_max;
  const int32_t *nq, *nk;
  const float* lse;
  float* colsum;  /* [n_img, Nk_max]; row = QUERY image img (sums over the keys of image src(img) = (img + src_offset) % n_img),
                     or row = the KEY image src(img) when by_key_image != 0 */
  const void *q_lo, *k_lo; /* optional, both or none */
  float* scratch;          /* [n_img, 4, Nk_max] workspace */
  int32_t by_key_image;
  int32_t _pad;
} imp_attn_colsum_args;
IMP_API int imp_attention_colsum(const imp_attn_colsum_args* args, void* stream);

/* ---- InstanceNorm1d(eps, affine=False) over tokens + ReLU, nets/layers.py:68-72 ----------------------------- */
/* H fp32 [batch, N, C] (row stride ldh) -> fp16 planes (or fp32 if out_f32 != NULL) with row stride ldo */
IMP_API int imp_instnorm_relu(const float* H, int64_t h_batch_stride, int32_t ldh, const int32_t* ns, int32_t Nmax, int32_t C,
                      int32_t batch, float eps, int32_t relu, void* out_hi, void* out_lo, float* out_f32,
                      int64_t o_batch_stride, int32_t ldo, void* stream);

/* second half of the fused instance norm: reduce the statistics imp_gemm left (fp64, fixed order) to mean / 1/sqrt(var+eps)
 * per (image, channel) in `stats` [images][C][2], then stream H -> relu((H - mean) * rstd) -> fp16 hi/lo planes.
 * H [images * Np, C] fp32 contiguous; Np >= 128. */
IMP_API int imp_instnorm_apply(const float* H, const float* stat_partial, const float* stat_straddle, const int32_t* ns,
                       int32_t Np, int32_t C, int32_t images, float eps, int32_t relu, float* stats, void* out_hi,
                       void* out_lo, void* stream);
/* (out_hi == NULL: only the reduction to `stats` runs -- the normalisation then happens inside the consuming imp_gemm,
 * see imp_gemm_args.a_f32) */

/* ---- keypoint encoder narrow layers (3->32, 32->64), nets/layers.py:80-90 ----------------------------------- */
IMP_API int imp_kenc_input(const float* norm_kpts, const float* scores, float* out_xyz4, int64_t tokens, void* stream);
IMP_API int imp_small_linear(const float* X, int32_t ldx, const float* W, const float* bias, float* Y, int32_t ldy,
                     int64_t rows, int32_t Cin, int32_t Cout, void* stream);

/* ---- Sinkhorn optimal transport + fused arg-max, nets/layers.py:27-46 (sink_algorithm / sinkhorn) ----------- */
typedef struct imp_sinkhorn_args {
  const float* dist; /* [batch, N0max, ldd], 16-byte aligned, ldd and batch stride multiples of 4 floats */
  int64_t dist_batch_stride;
  int32_t ldd;
  int32_t iters;
  const float* bin_score; /* device scalar (model.bin_score) */
  float* P;               /* [batch, N0max+1, ldp] workspace = output scores (row stride ldp, multiple of 4) */
  int64_t p_batch_stride;
  int32_t ldp;
  int32_t _pad;
  float* u;          /* [batch, N0max+1] */
  float* colbuf;     /* [3, batch, ldp] */
  float* row_max;    /* [batch, N0max] */
  int32_t* row_arg;  /* [batch, N0max] */
  uint64_t* col_key; /* [batch, N1max] packed column arg-max */
  float* row_mass;   /* [batch, N0max] or NULL: sum_j scores[i, :N1] (pooling, nets/adgm.py:476) */
  float* col_mass;   /* [batch, N1max] or NULL */
  const int32_t *n0s, *n1s; /* per-sample sizes or NULL */
  int32_t N0max, N1max, batch;
  int32_t write_scores; /* 1: store the final (p u) v into P; 0: only arg-max / masses are produced (P is scratch) */
  /* Compact storage of softmax(M) for the iteration sweeps (the HBM-bound part: one sweep per iteration).  With
   * storage != IMP_SK_STORE_F32 the iterations stream a 16- or 24-bit copy of p from q_store, the row softmax
   * statistics are kept in row_stats, and the final scores / arg-max are recomputed from dist in fp32, so only the
   * scaling vectors u, v carry the storage rounding (measured deviations: DESIGN.md section 2).  Problems small enough
   * for the shared-memory-resident kernel ignore it.  q_store: q_batch_stride bytes per matrix, at least
   * (N0max+1) * roundup16(N1max+1) * (2 or 3) bytes, 16-byte aligned.  row_stats: [2, batch, N0max+1] floats. */
  void* q_store;
  int64_t q_batch_stride;
  float* row_stats;
  int32_t storage; /* IMP_SK_STORE_* */
  int32_t _pad2;
} imp_sinkhorn_args;
#define IMP_SK_STORE_F32 0 /* exact fp32 copy of softmax(M): the reference recurrence on fp32 probabilities (default) */
#define IMP_SK_STORE_F16 1 /* p * 2^14 as IEEE fp16: 2 bytes per element */
#define IMP_SK_STORE_F24 2 /* top 16 bits of the fp32 word + one byte of mantissa extension (planar): 3 bytes per element */
/* flag OR-ed into `storage`: never take the shared-memory-resident kernel (a COOPERATIVE launch) for this problem -- for
 * callers that keep several Sinkhorn problems in flight on different streams, where cooperative grids could wait for each
 * other's resources */
#define IMP_SK_NO_RESIDENT 0x100
IMP_API int imp_sinkhorn(const imp_sinkhorn_args* args, void* stream);
/* Workspace query (the library never allocates): bytes PER MATRIX of q_store that imp_sinkhorn wants for this problem size
 * and storage format, or 0 when it will not use one (small problems run a shared-memory-resident kernel, N1 outside
 * [63, 4095] the row-ring kernels).  Needs a CUDA device (the answer depends on its SM count). */
IMP_API int64_t imp_sinkhorn_q_store_bytes(int32_t batch, int32_t N0max, int32_t N1max, int32_t storage);
/* Measurement aid for bench.py: with profiling on, imp_sinkhorn brackets its iteration launches with CUDA events on the
 * launching stream; imp_sinkhorn_iter_ms() waits for them and returns the mean duration of one iteration kernel of the
 * most recent call (< 0 if none was recorded). */
/* geometry query, no GPU needed: rows per work item of the streaming Sinkhorn kernels for `batch` matrices of N0max + 1 rows on a
 * persistent grid of `resident_ctas` CTAs (2 per SM).  Exposed for the regression test of the work-decomposition heuristic. */
IMP_API int imp_sinkhorn_rows_per_item(int32_t batch, int32_t N0max, int32_t resident_ctas);
IMP_API int imp_set_profiling(int32_t on);
IMP_API float imp_sinkhorn_iter_ms(void);

/* mutual-NN matches, GM.compute_matches nets/gm.py:305-320 (int64 indices like torch).  Every entry of the outputs is
 * written: rows / columns beyond a sample's n0s[b] / n1s[b] get the "no match" defaults (-1, 0). */
typedef struct imp_match_args {
  const float* row_max;
  const int32_t* row_arg;
  const uint64_t* col_key;
  float p_thresh;
  int32_t _pad;
  int64_t *indices0, *indices1;
  float *mscores0, *mscores1;
  const int32_t *n0s, *n1s;
  int32_t N0max, N1max, batch, _pad2;
  int64_t out0_batch_stride, out1_batch_stride;
} imp_match_args;
IMP_API int imp_matches(const imp_match_args* args, void* stream);

/* dual-softmax scorer, nets/layers.py:20-24 (with_sinkhorn=False).  n0s / n1s (device, [batch], may be NULL): sample b
 * scores its leading n0s[b] x n1s[b] block with the dustbin row / column right behind it (EIMP's kept subsets,
 * nets/adgm.py:441-447); the rest of P[b] is written as 0. */
IMP_API int imp_dual_softmax(const float* dist, int64_t dist_batch_stride, int32_t ldd, const float* bin_score, float* P,
                     int64_t p_batch_stride, int32_t ldp, float* row_lse, float* col_lse, int32_t N0, int32_t N1,
                     int32_t batch, const int32_t* n0s, const int32_t* n1s, void* stream);
/* row / column arg-max (and optional masses, may be NULL) of an existing score matrix over P[:, :N0, :N1]
 * (compute_matches / pool on caller-provided scores) */
IMP_API int imp_score_argmax(const float* P, int64_t p_batch_stride, int32_t ldp, float* row_max, int32_t* row_arg,
                     uint64_t* col_key, float* row_mass, float* col_mass, int32_t N0, int32_t N1, int32_t batch,
                     const int32_t* n0s, const int32_t* n1s, void* stream);

/* ---- EIMP adaptive pooling, nets/adgm.py:463-500 and :552-605 ------------------------------------------------
 * keep = { i : mass_i >= thresh } U { i : a_self_i >= lower_median(a_self[pids]) } U { i : a_cross_i >= ... },
 * evaluated on the current kept subset of cnt_in[b] tokens whose global ids are ids_in[b, 0..cnt) (sorted); writes
 * the new sorted global ids and count.  If cnt_in <= n_min_tokens or no row passes the threshold the subset is
 * copied unchanged and changed[b] = 0.  mass / a_self / a_cross are indexed by SUBSET POSITION; a_* are the
 * un-normalised received-attention sums (pruned tokens receive exactly 0 in the reference, so normalising by the
 * subset total equals nets/adgm.py:429-432). */
typedef struct imp_pool_args {
  const float* mass;             /* [batch, ld] */
  const float *a_self, *a_cross; /* [batch, ld] */
  int32_t ld;                    /* row stride of mass / a_self / a_cross / ids */
  int32_t _pad0;
  const int32_t* ids_in;  /* [batch, ld] */
  const int32_t* cnt_in;  /* [batch] */
  int32_t* ids_out;       /* [batch, ld] */
  int32_t* cnt_out;       /* [batch] */
  int32_t* changed;       /* [batch] */
  float thresh;
  int32_t n_min_tokens;
  int32_t batch;
  int32_t _pad;
} imp_pool_args;
IMP_API int imp_pool_select(const imp_pool_args* args, void* stream);

/* scatter subset-coordinate matches back to global ids (nets/adgm.py:458-460):
 *   out_idx[b, gids0[b, r]] = gids1[b, idx0[b, r]] if idx0[b, r] >= 0;  out_ms[b, gids0[b, r]] = ms0[b, r];  r < cnt0[b]
 * out_idx must be pre-filled with -1 and out_ms with 0. */
IMP_API int imp_scatter_matches(const int64_t* idx0, const float* ms0, int32_t ld_sub, const int32_t* gids0,
                        const int32_t* gids1, int32_t ld_ids, const int32_t* cnt0, int64_t* out_idx, float* out_ms,
                        int32_t ld_out, int32_t batch, void* stream);

/* row gather for compaction: out[b, r, :copy_bytes] = in[b, ids[b, r], :copy_bytes], r < cnt[b] */
IMP_API int imp_gather_rows(const void* in, int64_t in_batch_stride_bytes, int32_t row_bytes_in, const int32_t* ids,
                    int32_t ids_ld, const int32_t* cnt, void* out, int64_t out_batch_stride_bytes,
                    int32_t row_bytes_out, int32_t copy_bytes, int32_t max_rows, int32_t batch, void* stream);

/* ==== SuperPoint front-end (reference nets/superpoint.py:97-235) ===================================
 * The step before the matcher: image -> keypoints, scores, 256-d descriptors.  Activations are NHWC fp16 hi/lo planes
 * [B, H, W, C] (x = hi + lo, the matcher's split-precision format), weights are [C_out, 9 * C_in] hi/lo planes with
 * k = (3 * ky + kx) * C_in + ci, i.e. conv.weight.permute(0, 2, 3, 1). */

/* 3 x 3 convolution, stride 1, zero padding 1, + bias (+ ReLU): conv1b ... convPa / convDa (nets/superpoint.py:126-141,
 * 186-197).  tcgen05 implicit GEMM; C_in % 64 == 0, C_out in {64, 128, 256}. */
typedef struct imp_sp_conv_args {
  const void* in_hi;  /* fp16 [B, H, W, Cin] */
  const void* in_lo;
  const void* w_hi;   /* fp16 [Cout, 9 * Cin] */
  const void* w_lo;
  const float* bias;  /* [Cout] */
  void* out_hi;       /* fp16 [B, H, W, Cout] */
  void* out_lo;
  int32_t B, H, W, Cin, Cout, relu;
  int32_t pool;       /* 1: nn.MaxPool2d(2, 2) fused into the epilogue, out is [B, H/2, W/2, Cout] */
  int32_t _pad;
} imp_sp_conv_args;
IMP_API int imp_sp_conv3x3(const imp_sp_conv_args* args, void* stream);

/* conv1a (1 -> 64 channels, nets/superpoint.py:125) + ReLU: img fp32 [B, H, W], w fp32 [64, 9], out planes [B, H, W, 64] */
IMP_API int imp_sp_conv1a(const float* img, const float* w, const float* bias, void* out_hi, void* out_lo, int32_t B, int32_t H,
                          int32_t W, void* stream);
/* nn.MaxPool2d(2, 2) on planes (nets/superpoint.py:122): [B, H, W, C] -> [B, H/2, W/2, C] */
IMP_API int imp_sp_maxpool2(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int32_t B, int32_t H, int32_t W,
                            int32_t C, void* stream);
/* softmax over the 65 logits of a cell, dust bin dropped, depth-to-space (nets/superpoint.py:199-203):
 * logits fp32 [B * Hc * Wc, ld] (first 65 columns) -> scores fp32 [B, 8 Hc, 8 Wc] */
IMP_API int imp_sp_scores(const float* logits, int32_t ld, float* scores, int32_t B, int32_t Hc, int32_t Wc, void* stream);
/* simple_nms (nets/superpoint.py:50-66): mask[b, y, x] = 1 where the score survives; supp is a scratch of the same size */
IMP_API int imp_sp_nms(const float* scores, uint8_t* mask, uint8_t* supp, int32_t B, int32_t H, int32_t W, int32_t radius,
                       void* stream);

/* keypoints of ONE image (nets/superpoint.py:206-225): nonzero(nms > threshold) in row-major order, remove_borders, top-k by
 * score (all candidates in row-major order when there are at most max_keypoints, or max_keypoints < 0); (h, w) -> (x, y). */
typedef struct imp_sp_select_args {
  const float* scores;   /* [H, W] */
  const uint8_t* mask;   /* [H, W] from imp_sp_nms */
  int32_t H, W;
  float threshold;
  int32_t border;
  int32_t max_keypoints;
  int32_t cap;           /* capacity of the candidate arrays */
  int32_t* rowcnt;       /* scratch [H] */
  int32_t* rowoff;       /* scratch [H] */
  int32_t* total;        /* out: number of candidates */
  int32_t* cand_yx;      /* scratch [cap, 2] */
  float* cand_score;     /* scratch [cap] */
  uint64_t* keys;        /* scratch [next power of two >= cap] */
  float* kpts_xy;        /* out [min(cap, max_keypoints), 2]; rows behind the n_out keypoints are zeroed up to max_keypoints */
  float* kscores;        /* out */
  int32_t* n_out;        /* out: number of keypoints */
} imp_sp_select_args;
IMP_API int imp_sp_select(const imp_sp_select_args* args, void* stream);

/* x[r, :256] /= max(||x[r, :256]||, 1e-12) (F.normalize over the channels of the dense descriptor map, :229) */
IMP_API int imp_sp_l2norm_rows(float* x, int64_t rows, int32_t ld, void* stream);
/* sample_descriptors (nets/superpoint.py:83-95): bilinear interpolation of the normalised [Hc * Wc, 256] map at the keypoints
 * (grid_sample with its default align_corners=False -- the reference's version test `int(torch.__version__[2]) > 2` is false
 * for 1.12 and for 2.x alike -- zero padding), then a second L2 normalisation.  out fp32 [max_k, 256]: rows n = min(*n_kpts, max_k) .. max_k are zeroed (n_kpts NULL: n = max_k). */
IMP_API int imp_sp_sample_descriptors(const float* dmap, const float* kpts_xy, const int32_t* n_kpts, float* out, int32_t Hc,
                                      int32_t Wc, int32_t max_k, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IMP_B200_H_ */
