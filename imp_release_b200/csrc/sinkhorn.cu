// Sinkhorn optimal-transport assignment, probability domain, bit-for-bit the reference's recurrence
// (nets/layers.py:27-46): pad a dustbin column + row with bin_score, p = softmax_rows(M), u = v = 1,
// S x { u = r / (sum_j p v + 1e-8);  v = c / (sum_i p u + 1e-8) },  out = (p u) v,
// r = [1..1, N0+1], c = [1..1, N1+1].  Row softmax is the max-subtracted (LSE-stabilised) form.
//
// HBM-bound streaming kernels built around one "row ring": a producer warp streams whole matrix rows (8 KB at
// N = 2000) into a shared-memory ring with 1-D bulk TMA copies (cp.async.bulk + mbarrier complete_tx), ~100 KB in
// flight per CTA, two CTAs per SM; four consumer warps take rows off the ring.  One launch per Sinkhorn iteration
// reads the matrix ONCE: the consumer reduces the row against v (warp shuffles) to get u_i, then immediately folds
// p_ij * u_i into per-lane column accumulators; the CTA combines its warps in shared memory and issues one global
// atomicAdd per column.  v is never materialised: the next launch recomputes c_j / (colsum_j + eps) on the fly from
// the accumulated column sums (three rotating buffers: read / accumulate / being-zeroed).  The final launch applies
// (p u) v (written back only when the caller wants the score matrix) with the row arg-max and the row / column
// masses EIMP's pooling needs; a coalesced column pass produces the column arg-max (lowest index wins ties, like
// torch.max on CPU).
#include "sinkhorn.cuh"

#include <cooperative_groups.h>
#include <float.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"
#include "sinkhorn_common.cuh"

namespace imp {

static constexpr int SKR_CONSUMERS = 4;
static constexpr int SKR_THREADS = (SKR_CONSUMERS + 1) * 32;  // + one producer warp
static constexpr int SKR_SMEM_BUDGET = 113 * 1024;            // two CTAs per SM (228 KB - 2 x 1 KB reserved)

struct SkParams {
  const float* dist;
  long long dist_bs;
  int ldd;
  const float* bin_score;
  float* P;
  long long p_bs;
  int ldp;
  float* u;
  const float* col_prev;
  float* col_acc;
  float* col_zero;
  float* row_max;
  int* row_arg;
  float* row_mass;
  float* col_mass;
  const int *n0s, *n1s;
  int N0max, N1max;
  int rows_per_cta, ring_slots;
  int do_iter;       // init: also perform the first half-iteration;  final: Sinkhorn ran at least once
  int write_scores;  // final: store (p u) v back into P
};

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(SKR_CONSUMERS * 32) : "memory"); }

enum { SK_INIT = 0, SK_ITER = 1, SK_FINAL = 2 };

// MODE: SK_INIT  P = softmax_rows(pad(dist)) (+ first half-iteration: u with v = 1, column sums with that u)
//       SK_ITER  one full Sinkhorn iteration (u, then column sums) in a single sweep over P
//       SK_FINAL out = (p u) v, row arg-max / masses over the non-dustbin block
template <int NV, int MODE>
__global__ void __launch_bounds__(SKR_THREADS, 2) sk_ring_kernel(const SkParams p) {
  extern __shared__ __align__(16) float sk_smem[];
  const int b = blockIdx.y;
  const SkDims d = sk_dims(p.n0s, p.n1s, b, p.N0max, p.N1max);
  const int row0 = blockIdx.x * p.rows_per_cta;
  if (row0 >= d.R) return;
  const int nrows = min(p.rows_per_cta, d.R - row0);
  const int S = p.ring_slots;
  float* ring = sk_smem;                         // [S][ldp]
  float* s_v = ring + (size_t)S * p.ldp;         // [ldp]
  float* s_col = s_v + p.ldp;                    // [ldp]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_col + p.ldp);
  uint64_t* empty_bar = full_bar + S;
  const int warp = threadIdx.x >> 5;
  const int C4 = (d.C + 3) & ~3;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == SKR_CONSUMERS) {
    // ------------------------------------------------------------------ producer: stream rows into the ring
    if (lane_id() == 0) {
      const float* src = (MODE == SK_INIT) ? p.dist + b * p.dist_bs : p.P + b * p.p_bs;
      const long long ld = (MODE == SK_INIT) ? p.ldd : p.ldp;
      const uint32_t bytes = (MODE == SK_INIT) ? (uint32_t)(((d.C - 1 + 3) & ~3) * 4) : (uint32_t)(C4 * 4);
      for (int r = 0; r < nrows; ++r) {
        const int s = r % S;
        mbar_wait(&empty_bar[s], ((r / S) & 1) ^ 1);
        const int i = row0 + r;
        if (MODE == SK_INIT && (i == d.R - 1 || bytes == 0)) {
          mbar_arrive(&full_bar[s]);  // the dustbin row has no source: the consumer synthesises it
        } else {
          mbar_arrive_expect_tx(&full_bar[s], bytes);
          bulk_copy_g2s(ring + (size_t)s * p.ldp, src + (long long)i * ld, bytes, &full_bar[s]);
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const int ct = threadIdx.x;  // 0..127
  if (blockIdx.x == 0 && p.col_zero != nullptr)
    for (int j = ct; j < p.ldp; j += SKR_CONSUMERS * 32) p.col_zero[(long long)b * p.ldp + j] = 0.f;
  for (int c0 = 4 * ct; c0 < C4; c0 += 4 * SKR_CONSUMERS * 32) {
    float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
    if (MODE == SK_ITER || (MODE == SK_FINAL && p.do_iter)) v = v_from_colsum(p.col_prev + (long long)b * p.ldp, c0, d.C);
    *reinterpret_cast<float4*>(s_v + c0) = v;
    *reinterpret_cast<float4*>(s_col + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  consumer_sync();

  const float bin = (MODE == SK_INIT) ? *p.bin_score : 0.f;
  const bool want_col = (MODE != SK_FINAL) ? (p.do_iter != 0 || MODE == SK_ITER) : (p.col_mass != nullptr);
  float4 acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int r = warp; r < nrows; r += SKR_CONSUMERS) {
    const int s = r % S;
    const int i = row0 + r;
    mbar_wait(&full_bar[s], (r / S) & 1);
    float* srow = ring + (size_t)s * p.ldp;
    // Rows are re-read from the ring instead of being held in registers (keeps the kernel at ~2 CTAs/SM without
    // spills); each lane only ever touches its own float4 groups, so in-place updates of the slot are race-free.
    if (MODE == SK_INIT) {
      const bool bin_row = (i == d.R - 1);
      // pass 1: materialise the padded row in the slot (dustbin column / row, -FLT_MAX beyond C) and take its max.
      // Interior float4 groups (all four columns < C-1) need no per-element masking.
      float m = -FLT_MAX;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c0 = 4 * (lane_id() + 32 * k);
        if (c0 < d.C) {
          float4 t;
          if (!bin_row && c0 + 3 < d.C - 1) {
            t = *reinterpret_cast<const float4*>(srow + c0);
          } else {
            t = make_float4(bin, bin, bin, bin);
            if (!bin_row && c0 < d.C - 1) t = *reinterpret_cast<const float4*>(srow + c0);
            t.x = (c0 + 0 < d.C) ? ((bin_row || c0 + 0 == d.C - 1) ? bin : t.x) : -FLT_MAX;
            t.y = (c0 + 1 < d.C) ? ((bin_row || c0 + 1 == d.C - 1) ? bin : t.y) : -FLT_MAX;
            t.z = (c0 + 2 < d.C) ? ((bin_row || c0 + 2 == d.C - 1) ? bin : t.z) : -FLT_MAX;
            t.w = (c0 + 3 < d.C) ? ((bin_row || c0 + 3 == d.C - 1) ? bin : t.w) : -FLT_MAX;
            *reinterpret_cast<float4*>(srow + c0) = t;
          }
          m = fmaxf(m, fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)));
        }
      }
      m = warp_max(m);
      // pass 2: e = exp(x - max) in place, row sum  (sk_exp(-FLT_MAX - m) underflows to exactly 0 for the pad columns)
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c0 = 4 * (lane_id() + 32 * k);
        if (c0 < d.C) {
          float4 t = *reinterpret_cast<const float4*>(srow + c0);
          t.x = sk_exp(t.x - m);
          t.y = sk_exp(t.y - m);
          t.z = sk_exp(t.z - m);
          t.w = sk_exp(t.w - m);
          *reinterpret_cast<float4*>(srow + c0) = t;
          sum += (t.x + t.y) + (t.z + t.w);
        }
      }
      sum = warp_sum(sum);
      const float inv_sum = 1.0f / sum;  // one division per row; p = e * (1/sum) differs from e / sum by <= 1 ulp
      // first half-iteration with v = 1: sum_j p_ij = sum * inv_sum (the reference adds the rounded p's; the two agree
      // to ~1e-7 relative)
      const float ui = p.do_iter ? (bin_row ? (float)d.R : 1.f) / (sum * inv_sum + SK_EPS) : 0.f;
      if (p.do_iter && lane_id() == 0) p.u[(long long)b * (p.N0max + 1) + i] = ui;
      // pass 3: normalise, store, and fold p * u into the column accumulators
      float* prow = p.P + b * p.p_bs + (long long)i * p.ldp;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c0 = 4 * (lane_id() + 32 * k);
        if (c0 < d.C) {
          float4 t = *reinterpret_cast<const float4*>(srow + c0);
          t.x *= inv_sum;
          t.y *= inv_sum;
          t.z *= inv_sum;
          t.w *= inv_sum;
          *reinterpret_cast<float4*>(prow + c0) = t;
          acc[k].x += t.x * ui;
          acc[k].y += t.y * ui;
          acc[k].z += t.z * ui;
          acc[k].w += t.w * ui;
        } else if (c0 < p.ldp) {
          *reinterpret_cast<float4*>(prow + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else if (MODE == SK_ITER) {
      float rs = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c0 = 4 * (lane_id() + 32 * k);
        if (c0 < d.C) {
          const float4 t = *reinterpret_cast<const float4*>(srow + c0);
          const float4 v = *reinterpret_cast<const float4*>(s_v + c0);
          rs += (t.x * v.x + t.y * v.y) + (t.z * v.z + t.w * v.w);
        }
      }
      rs = warp_sum(rs);
      const float ui = ((i == d.R - 1) ? (float)d.R : 1.f) / (rs + SK_EPS);
      if (lane_id() == 0) p.u[(long long)b * (p.N0max + 1) + i] = ui;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c0 = 4 * (lane_id() + 32 * k);
        if (c0 < d.C) {
          const float4 t = *reinterpret_cast<const float4*>(srow + c0);
          acc[k].x += t.x * ui;
          acc[k].y += t.y * ui;
          acc[k].z += t.z * ui;
          acc[k].w += t.w * ui;
        }
      }
    } else {  // SK_FINAL
      const float ui = p.do_iter ? p.u[(long long)b * (p.N0max + 1) + i] : 1.f;
      const bool inner_row = i < d.R - 1;
      float* prow = p.P + b * p.p_bs + (long long)i * p.ldp;
      // four independent (value, column) trackers -- one per float4 component -- keep the compare/select chains short
      float bv[4] = {-1.f, -1.f, -1.f, -1.f}, ms[4] = {0.f, 0.f, 0.f, 0.f};
      int bj[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c0 = 4 * (lane_id() + 32 * k);
        if (c0 >= d.C) continue;
        const float4 t = *reinterpret_cast<const float4*>(srow + c0);
        const float4 v = *reinterpret_cast<const float4*>(s_v + c0);
        const float o[4] = {(t.x * ui) * v.x, (t.y * ui) * v.y, (t.z * ui) * v.z, (t.w * ui) * v.w};
        if (p.write_scores) *reinterpret_cast<float4*>(prow + c0) = make_float4(o[0], o[1], o[2], o[3]);
        if (inner_row) {
          if (c0 + 3 < d.C - 1) {  // interior group: no column masking needed
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              ms[q] += o[q];
              if (o[q] > bv[q]) {
                bv[q] = o[q];
                bj[q] = c0 + q;
              }
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const bool in = c0 + q < d.C - 1;
              const float oq = in ? o[q] : -1.f;  // columns grow with k, so a strict > keeps the lowest column per tracker
              ms[q] += in ? o[q] : 0.f;
              if (oq > bv[q]) {
                bv[q] = oq;
                bj[q] = c0 + q;
              }
            }
          }
          if (want_col) {
            acc[k].x += o[0];
            acc[k].y += o[1];
            acc[k].z += o[2];
            acc[k].w += o[3];
          }
        }
      }
      float best = bv[0], mass = (ms[0] + ms[1]) + (ms[2] + ms[3]);
      int best_j = bj[0];
#pragma unroll
      for (int q = 1; q < 4; ++q)
        if (bv[q] > best || (bv[q] == best && bj[q] < best_j)) {
          best = bv[q];
          best_j = bj[q];
        }
      if (inner_row) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {  // warp arg-max, lowest index wins ties
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
          if (ob > best || (ob == best && oj < best_j)) {
            best = ob;
            best_j = oj;
          }
        }
        mass = warp_sum(mass);
        if (lane_id() == 0) {
          p.row_max[(long long)b * p.N0max + i] = best;
          p.row_arg[(long long)b * p.N0max + i] = best_j;
          if (p.row_mass) p.row_mass[(long long)b * p.N0max + i] = mass;
        }
      }
    }
    __syncwarp();
    if (lane_id() == 0) mbar_arrive(&empty_bar[s]);  // the ring slot is free again
  }

  if (want_col) {
    // combine the CTA's warps in shared memory, then one global atomic per column
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = 4 * (lane_id() + 32 * k);
      if (c0 < d.C) {
        atomicAdd(s_col + c0 + 0, acc[k].x);
        atomicAdd(s_col + c0 + 1, acc[k].y);
        atomicAdd(s_col + c0 + 2, acc[k].z);
        atomicAdd(s_col + c0 + 3, acc[k].w);
      }
    }
    consumer_sync();
    if (MODE == SK_FINAL) {
      for (int j = ct; j < d.C - 1; j += SKR_CONSUMERS * 32) atomicAdd(p.col_mass + (long long)b * p.N1max + j, s_col[j]);
    } else {
      for (int j = ct; j < d.C; j += SKR_CONSUMERS * 32) atomicAdd(p.col_acc + (long long)b * p.ldp + j, s_col[j]);
    }
  }
}

// column arg-max over the non-dustbin block: thread per column (coalesced), 256-row slabs, packed atomicMax
__global__ void __launch_bounds__(128)
sk_colmax_kernel(const float* __restrict__ P, long long p_bs, int ldp, const float* __restrict__ u,
                 const float* __restrict__ col_last, int scaled, int has_iter, unsigned long long* __restrict__ col_key,
                 const int* __restrict__ n0s, const int* __restrict__ n1s, int N0max, int N1max, int slab) {
  const int b = blockIdx.z;
  const SkDims d = sk_dims(n0s, n1s, b, N0max, N1max);
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i0 = blockIdx.y * slab;
  if (j >= d.C - 1 || i0 >= d.R - 1) return;
  const int i1 = min(i0 + slab, d.R - 1);
  float vj = 1.f;
  if (!scaled && has_iter) vj = 1.f / (col_last[(long long)b * ldp + j] + SK_EPS);  // c_j = 1 for j < C-1
  const float* base = P + b * p_bs + j;
  const float* ub = u + (long long)b * (N0max + 1);
  float best = -1.f;
  int bi = 0;
#pragma unroll 4
  for (int i = i0; i < i1; ++i) {
    float val = base[(long long)i * ldp];
    if (!scaled) val = (val * (has_iter ? ub[i] : 1.f)) * vj;
    if (val > best) {
      best = val;
      bi = i;
    }
  }
  atomicMax(col_key + (long long)b * N1max + j, pack_max_key(best, bi));
}

// mutual nearest-neighbour check + threshold (GM.compute_matches, nets/gm.py:305-320)
__global__ void sk_matches_kernel(const float* __restrict__ row_max, const int* __restrict__ row_arg,
                                  const unsigned long long* __restrict__ col_key, float p_thresh,
                                  long long* __restrict__ indices0, long long* __restrict__ indices1,
                                  float* __restrict__ mscores0, float* __restrict__ mscores1,
                                  const int* __restrict__ n0s, const int* __restrict__ n1s, int N0max, int N1max,
                                  long long out0_bs, long long out1_bs) {
  const int b = blockIdx.y;
  const int n0 = n0s ? n0s[b] : N0max;
  const int n1 = n1s ? n1s[b] : N1max;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long* ck = col_key + (long long)b * N1max;
  const int* ra = row_arg + (long long)b * N0max;
  const float* rm = row_max + (long long)b * N0max;
  if (t < n0) {
    const int j = ra[t];
    const int back = (int)(0xFFFFFFFFu - (unsigned)(ck[j] & 0xFFFFFFFFull));
    const bool mutual = back == t;
    const float ms = mutual ? rm[t] : 0.f;
    mscores0[b * out0_bs + t] = ms;
    indices0[b * out0_bs + t] = (mutual && ms > p_thresh) ? (long long)j : -1ll;
  } else if (t < N0max) {  // beyond this sample's keypoints: the "no match" defaults (outputs need no pre-fill)
    mscores0[b * out0_bs + t] = 0.f;
    indices0[b * out0_bs + t] = -1ll;
  }
  if (t >= n1 && t < N1max && indices1 != nullptr) {
    mscores1[b * out1_bs + t] = 0.f;
    indices1[b * out1_bs + t] = -1ll;
  }
  if (t < n1 && indices1 != nullptr) {
    const int i = (int)(0xFFFFFFFFu - (unsigned)(ck[t] & 0xFFFFFFFFull));
    const bool mutual1 = ra[i] == t;  // then row i is mutual too and mscores0[i] = row_max[i]
    const float ms0_i = mutual1 ? rm[i] : 0.f;
    mscores1[b * out1_bs + t] = ms0_i;
    indices1[b * out1_bs + t] = (mutual1 && ms0_i > p_thresh) ? (long long)i : -1ll;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Resident variant for small batches (one pair at N ~ 2000: the B = 1 latency path and the 2048^2 x 100-iteration
// micro-benchmark): the whole matrix (16 MB) fits in the aggregate shared memory of one cooperative grid (2 CTAs/SM x
// 8 rows x 8 KB), so it is read from HBM exactly once, all Sinkhorn iterations run out of shared memory inside ONE
// launch with a grid-wide barrier per iteration, and only the final scores are written back.
namespace cg = cooperative_groups;
static constexpr int SKS_THREADS = 128;
static constexpr int SKS_WARPS = 4;

template <int NV>
__global__ void __launch_bounds__(SKS_THREADS, 2) sk_resident_kernel(const SkParams p, float* col0, float* col1, float* col2,
                                                                     unsigned long long* col_key, int iters) {
  extern __shared__ __align__(16) float sk_smem[];
  cg::grid_group grid = cg::this_grid();
  const int ctas_per_mat = (p.N0max + 1 + p.rows_per_cta - 1) / p.rows_per_cta;
  const int b = blockIdx.x / ctas_per_mat;
  const SkDims d = sk_dims(p.n0s, p.n1s, b, p.N0max, p.N1max);
  const int row0 = (blockIdx.x % ctas_per_mat) * p.rows_per_cta;
  const int nrows = max(0, min(p.rows_per_cta, d.R - row0));
  float* rows = sk_smem;                                  // [rows_per_cta][ldp]
  float* s_v = rows + (size_t)p.rows_per_cta * p.ldp;     // [ldp]
  float* s_col = s_v + p.ldp;                             // [ldp]
  float* s_u = s_col + p.ldp;                             // [rows_per_cta]
  const int warp = threadIdx.x >> 5;
  const int C4 = (d.C + 3) & ~3;
  const float bin = *p.bin_score;
  float* cols[3] = {col0 + (long long)b * p.ldp, col1 + (long long)b * p.ldp, col2 + (long long)b * p.ldp};

  // ---- load + softmax (same arithmetic as the streaming init pass)
  for (int r = warp; r < nrows; r += SKS_WARPS) {
    const int i = row0 + r;
    const bool bin_row = (i == d.R - 1);
    float* srow = rows + (size_t)r * p.ldp;
    const float* drow = p.dist + b * p.dist_bs + (long long)i * p.ldd;
    float m = -FLT_MAX;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = 4 * (lane_id() + 32 * k);
      if (c0 < C4) {
        float4 t = make_float4(bin, bin, bin, bin);
        if (!bin_row && c0 < d.C - 1) t = *reinterpret_cast<const float4*>(drow + c0);  // ldd >= roundup4(N1)
        t.x = (c0 + 0 < d.C) ? ((bin_row || c0 + 0 == d.C - 1) ? bin : t.x) : -FLT_MAX;
        t.y = (c0 + 1 < d.C) ? ((bin_row || c0 + 1 == d.C - 1) ? bin : t.y) : -FLT_MAX;
        t.z = (c0 + 2 < d.C) ? ((bin_row || c0 + 2 == d.C - 1) ? bin : t.z) : -FLT_MAX;
        t.w = (c0 + 3 < d.C) ? ((bin_row || c0 + 3 == d.C - 1) ? bin : t.w) : -FLT_MAX;
        *reinterpret_cast<float4*>(srow + c0) = t;
        m = fmaxf(m, fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)));
      }
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = 4 * (lane_id() + 32 * k);
      if (c0 < C4) {
        float4 t = *reinterpret_cast<const float4*>(srow + c0);
        t.x = sk_exp(t.x - m);
        t.y = sk_exp(t.y - m);
        t.z = sk_exp(t.z - m);
        t.w = sk_exp(t.w - m);
        *reinterpret_cast<float4*>(srow + c0) = t;
        sum += (t.x + t.y) + (t.z + t.w);
      }
    }
    sum = warp_sum(sum);
    const float inv_sum = 1.0f / sum;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = 4 * (lane_id() + 32 * k);
      if (c0 < C4) {
        float4 t = *reinterpret_cast<const float4*>(srow + c0);
        t.x *= inv_sum;
        t.y *= inv_sum;
        t.z *= inv_sum;
        t.w *= inv_sum;
        *reinterpret_cast<float4*>(srow + c0) = t;
      }
    }
    if (lane_id() == 0) s_u[r] = 1.f;
  }
  __syncthreads();

  // ---- Sinkhorn iterations out of shared memory
  for (int it = 0; it < iters; ++it) {
    const float* prev = cols[(it + 2) % 3];
    float* accg = cols[it % 3];
    float* zero = cols[(it + 1) % 3];
    if (blockIdx.x % ctas_per_mat == 0)
      for (int j = threadIdx.x; j < p.ldp; j += SKS_THREADS) zero[j] = 0.f;
    for (int c0 = 4 * threadIdx.x; c0 < C4; c0 += 4 * SKS_THREADS) {
      *reinterpret_cast<float4*>(s_v + c0) = (it == 0) ? make_float4(c0 + 0 < d.C ? 1.f : 0.f, c0 + 1 < d.C ? 1.f : 0.f,
                                                                     c0 + 2 < d.C ? 1.f : 0.f, c0 + 3 < d.C ? 1.f : 0.f)
                                                       : v_from_colsum(prev, c0, d.C);
      *reinterpret_cast<float4*>(s_col + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    float4 acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = warp; r < nrows; r += SKS_WARPS) {
      const int i = row0 + r;
      const float* srow = rows + (size_t)r * p.ldp;
      float rs = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c0 = 4 * (lane_id() + 32 * k);
        if (c0 < d.C) {
          const float4 t = *reinterpret_cast<const float4*>(srow + c0);
          const float4 v = *reinterpret_cast<const float4*>(s_v + c0);
          rs += (t.x * v.x + t.y * v.y) + (t.z * v.z + t.w * v.w);
        }
      }
      rs = warp_sum(rs);
      const float ui = ((i == d.R - 1) ? (float)d.R : 1.f) / (rs + SK_EPS);
      if (lane_id() == 0) s_u[r] = ui;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c0 = 4 * (lane_id() + 32 * k);
        if (c0 < d.C) {
          const float4 t = *reinterpret_cast<const float4*>(srow + c0);
          acc[k].x += t.x * ui;
          acc[k].y += t.y * ui;
          acc[k].z += t.z * ui;
          acc[k].w += t.w * ui;
        }
      }
    }
    if (nrows > 0) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c0 = 4 * (lane_id() + 32 * k);
        if (c0 < d.C && warp < nrows) {
          atomicAdd(s_col + c0 + 0, acc[k].x);
          atomicAdd(s_col + c0 + 1, acc[k].y);
          atomicAdd(s_col + c0 + 2, acc[k].z);
          atomicAdd(s_col + c0 + 3, acc[k].w);
        }
      }
    }
    __syncthreads();
    if (nrows > 0)
      for (int j = threadIdx.x; j < d.C; j += SKS_THREADS) atomicAdd(accg + j, s_col[j]);
    grid.sync();
  }

  // ---- final scaling, row arg-max / masses, column arg-max, optional write-back
  const float* last = cols[(iters + 2) % 3];
  for (int c0 = 4 * threadIdx.x; c0 < C4; c0 += 4 * SKS_THREADS)
    *reinterpret_cast<float4*>(s_v + c0) = iters > 0 ? v_from_colsum(last, c0, d.C)
                                                     : make_float4(c0 + 0 < d.C ? 1.f : 0.f, c0 + 1 < d.C ? 1.f : 0.f,
                                                                   c0 + 2 < d.C ? 1.f : 0.f, c0 + 3 < d.C ? 1.f : 0.f);
  __syncthreads();
  for (int r = warp; r < nrows; r += SKS_WARPS) {
    const int i = row0 + r;
    const float* srow = rows + (size_t)r * p.ldp;
    const float ui = iters > 0 ? s_u[r] : 1.f;
    const bool inner_row = i < d.R - 1;
    float* prow = p.P + b * p.p_bs + (long long)i * p.ldp;
    if (lane_id() == 0) p.u[(long long)b * (p.N0max + 1) + i] = ui;
    float best = -1.f, mass = 0.f;
    int best_j = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = 4 * (lane_id() + 32 * k);
      if (c0 >= d.C) {
        if (c0 < p.ldp && p.write_scores) *reinterpret_cast<float4*>(prow + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
        continue;
      }
      const float4 t = *reinterpret_cast<const float4*>(srow + c0);
      const float4 v = *reinterpret_cast<const float4*>(s_v + c0);
      const float o[4] = {(t.x * ui) * v.x, (t.y * ui) * v.y, (t.z * ui) * v.z, (t.w * ui) * v.w};
      // the resident path never materialised softmax(M) in HBM: P receives either the final scores or the plain p
      *reinterpret_cast<float4*>(prow + c0) = p.write_scores ? make_float4(o[0], o[1], o[2], o[3]) : t;
      if (inner_row) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (c0 + q < d.C - 1) {
            mass += o[q];
            if (o[q] > best) {
              best = o[q];
              best_j = c0 + q;
            }
            atomicMax(col_key + (long long)b * p.N1max + c0 + q,
                      (static_cast<unsigned long long>(__float_as_uint(o[q])) << 32) | (0xFFFFFFFFu - (unsigned)i));
            if (p.col_mass) atomicAdd(p.col_mass + (long long)b * p.N1max + c0 + q, o[q]);
          }
        }
      }
    }
    if (inner_row) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
        if (ob > best || (ob == best && oj < best_j)) {
          best = ob;
          best_j = oj;
        }
      }
      mass = warp_sum(mass);
      if (lane_id() == 0) {
        p.row_max[(long long)b * p.N0max + i] = best;
        p.row_arg[(long long)b * p.N0max + i] = best_j;
        if (p.row_mass) p.row_mass[(long long)b * p.N0max + i] = mass;
      }
    }
  }
}

// Optional L2-resident chunking (IMP_SK_L2_MB=<MB>): cut a big batch into chunks that fit in the 126 MB L2 and run the
// whole 22-launch sequence per chunk.  Measured on B200 at B=64, N=2000: 5.29 ms streaming vs 9.5 / 7.1 / 6.6 ms with
// 48 / 80 / 100 MB chunks (small launches + per-launch overhead lose more than the L2 hits win), so it is OFF by default.
static int sk_chunk_matrices(size_t matrix_bytes, int batch) {
  static long budget_mb = -1;
  if (budget_mb < 0) {
    const char* e = getenv("IMP_SK_L2_MB");
    budget_mb = e ? atol(e) : 0;  // measured on B200 (tools/sk_sweep.sh): chunking to 48/80/100 MB is SLOWER than streaming
  }
  if (budget_mb == 0) return batch;  // chunking disabled
  long c = (long)((size_t)budget_mb * 1024 * 1024 / matrix_bytes);
  if (c < 1) c = 1;
  return c > batch ? batch : (int)c;
}

// Will a problem of this size run in the shared-memory-resident kernel (which needs no q_store workspace)?  One source of
// truth for run_sinkhorn and for the workspace query of the C ABI.
static bool sk_resident_geometry(int batch, int R, int ldp, int* rows_per_cta, int* ctas_per_mat, size_t* smem) {
  const long long wave = 2LL * num_sms();
  int rpc = (int)(((long long)batch * R + wave - 1) / wave);
  rpc = (rpc + SKS_WARPS - 1) / SKS_WARPS * SKS_WARPS;
  const int cpm = (R + rpc - 1) / rpc;
  const size_t sm = ((size_t)rpc + 2) * (size_t)ldp * sizeof(float) + (size_t)rpc * sizeof(float) + 16;
  if (rows_per_cta) *rows_per_cta = rpc;
  if (ctas_per_mat) *ctas_per_mat = cpm;
  if (smem) *smem = sm;
  return (long long)batch * cpm <= wave && sm <= (size_t)SKR_SMEM_BUDGET;
}

// Run-time switch for the shared-memory-resident kernel (default on; IMP_SK_RESIDENT=0 or imp_set_option turn it off so
// that small problems take the streaming kernels too -- the parity tests use this to put the model-level fixtures
// through the path the big batches run).
static int g_sk_resident = -1;
static bool sk_resident_enabled() {
  if (g_sk_resident < 0) {
    const char* e = getenv("IMP_SK_RESIDENT");
    g_sk_resident = e ? (atoi(e) != 0) : 1;
  }
  return g_sk_resident != 0;
}
void sinkhorn_set_resident(int on) { g_sk_resident = on ? 1 : 0; }

long long sinkhorn_q_store_bytes(int batch, int N0max, int N1max, int storage) {
  const int R = N0max + 1, C = N1max + 1;
  if (batch <= 0 || N0max <= 0 || N1max <= 0 || C < 64 || C > 4096) return 0;
  const bool allow_resident = sk_resident_enabled() && !(storage & IMP_SK_NO_RESIDENT);
  storage &= ~IMP_SK_NO_RESIDENT;
  if (allow_resident && sk_resident_geometry(batch, R, (C + 3) & ~3, nullptr, nullptr, nullptr)) return 0;
  const int bpe = storage == IMP_SK_STORE_F16 ? 2 : (storage == IMP_SK_STORE_F24 ? 3 : 4);
  return (long long)R * ((C + 15) & ~15) * bpe;
}

static int g_profile = 0;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static int g_ev_iters = 0;
void sinkhorn_set_profiling(int on) { g_profile = on; }
float sinkhorn_iter_ms() {
  if (g_ev_iters <= 0 || g_ev0 == nullptr) return -1.f;
  if (cudaEventSynchronize(g_ev1) != cudaSuccess) return -1.f;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, g_ev0, g_ev1) != cudaSuccess) return -1.f;
  return ms / (float)g_ev_iters;
}
bool sk_profiling_on() { return g_profile != 0; }
void sk_profile_begin(cudaStream_t st) {
  if (g_ev0 == nullptr) {
    cudaEventCreate(&g_ev0);
    cudaEventCreate(&g_ev1);
  }
  cudaEventRecord(g_ev0, st);
}
void sk_profile_end(cudaStream_t st, int launches) {
  cudaEventRecord(g_ev1, st);
  g_ev_iters = launches;
}

template <int NV>
static int run_sinkhorn(const SinkhornArgs& a, cudaStream_t st) {
  const int R = a.N0max + 1;
  const size_t row_bytes = (size_t)a.ldp * sizeof(float);
  const size_t fixed = 2 * row_bytes + 2 * 64 * sizeof(uint64_t);
  static DeviceOnce configured;
  if (configured.first()) {
    IMP_CUDA_OK(cudaFuncSetAttribute(sk_ring_kernel<NV, SK_INIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKR_SMEM_BUDGET));
    IMP_CUDA_OK(cudaFuncSetAttribute(sk_ring_kernel<NV, SK_ITER>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKR_SMEM_BUDGET));
    IMP_CUDA_OK(cudaFuncSetAttribute(sk_ring_kernel<NV, SK_FINAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKR_SMEM_BUDGET));
    // two CTAs per SM need the full shared-memory carve-out
    IMP_CUDA_OK(cudaFuncSetAttribute(sk_ring_kernel<NV, SK_INIT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    IMP_CUDA_OK(cudaFuncSetAttribute(sk_ring_kernel<NV, SK_ITER>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    IMP_CUDA_OK(cudaFuncSetAttribute(sk_ring_kernel<NV, SK_FINAL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  const int chunk = sk_chunk_matrices((size_t)R * row_bytes, a.batch);
  const int iters = a.iters;
  IMP_CUDA_OK(cudaMemsetAsync(a.col_key, 0, (size_t)a.batch * a.N1max * sizeof(unsigned long long), st));
  if (a.col_mass) IMP_CUDA_OK(cudaMemsetAsync(a.col_mass, 0, (size_t)a.batch * a.N1max * sizeof(float), st));
  IMP_CUDA_OK(cudaMemsetAsync(a.colbuf, 0, (size_t)a.batch * a.ldp * sizeof(float), st));  // col[0]

  {  // small problems: everything resident in shared memory, one cooperative launch
    static bool coop_ok = true;  // cleared when a cooperative launch is refused (profiler / MPS)
    const bool use_resident = sk_resident_enabled() && coop_ok && !(a.storage & IMP_SK_NO_RESIDENT);
    int rpc, ctas_per_mat;
    size_t smem_res;
    const bool fits = sk_resident_geometry(a.batch, R, a.ldp, &rpc, &ctas_per_mat, &smem_res);
    if (use_resident && fits) {
      static DeviceOnce conf;
      if (conf.first()) {
        IMP_CUDA_OK(cudaFuncSetAttribute(sk_resident_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKR_SMEM_BUDGET));
        IMP_CUDA_OK(cudaFuncSetAttribute(sk_resident_kernel<NV>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      }
      int max_blocks = 0;
      IMP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, sk_resident_kernel<NV>, SKS_THREADS, smem_res));
      if ((long long)max_blocks * num_sms() >= (long long)a.batch * ctas_per_mat) {
        SkParams p;
        p.dist = a.dist; p.dist_bs = a.dist_batch_stride; p.ldd = a.ldd; p.bin_score = a.bin_score;
        p.P = a.P; p.p_bs = a.p_batch_stride; p.ldp = a.ldp; p.u = a.u;
        p.col_prev = nullptr; p.col_acc = nullptr; p.col_zero = nullptr;
        p.row_max = a.row_max; p.row_arg = a.row_arg; p.row_mass = a.row_mass; p.col_mass = a.col_mass;
        p.n0s = a.n0s; p.n1s = a.n1s; p.N0max = a.N0max; p.N1max = a.N1max;
        p.rows_per_cta = rpc; p.ring_slots = 0; p.do_iter = iters > 0; p.write_scores = a.write_scores;
        float* c0 = a.colbuf;
        float* c1 = a.colbuf + (size_t)a.batch * a.ldp;
        float* c2 = a.colbuf + 2 * (size_t)a.batch * a.ldp;
        unsigned long long* ck = reinterpret_cast<unsigned long long*>(a.col_key);
        int it = iters;
        void* args[] = {&p, &c0, &c1, &c2, &ck, &it};
        if (cudaLaunchCooperativeKernel((void*)sk_resident_kernel<NV>, dim3(a.batch * ctas_per_mat), dim3(SKS_THREADS), args,
                                        smem_res, st) == cudaSuccess)
          return 0;
        (void)cudaGetLastError();  // cooperative launch unavailable (e.g. under a profiler / MPS): use the streaming path
        coop_ok = false;
      }
    }
  }

  // big batches with a q_store workspace: column-split kernels over a stored (fp32 / 24-bit / fp16) copy of softmax(M)
  if (a.q_store != nullptr && a.row_stats != nullptr && a.N1max + 1 >= 64 && a.N1max + 1 <= 4096) return run_sinkhorn_compact(a, st);

  for (int b0 = 0; b0 < a.batch; b0 += chunk) {
    const int nb = (a.batch - b0 < chunk) ? a.batch - b0 : chunk;
    SkParams p;
    p.dist = a.dist + (long long)b0 * a.dist_batch_stride;
    p.dist_bs = a.dist_batch_stride;
    p.ldd = a.ldd;
    p.bin_score = a.bin_score;
    p.P = a.P + (long long)b0 * a.p_batch_stride;
    p.p_bs = a.p_batch_stride;
    p.ldp = a.ldp;
    p.u = a.u + (long long)b0 * (a.N0max + 1);
    p.row_max = a.row_max + (long long)b0 * a.N0max;
    p.row_arg = a.row_arg + (long long)b0 * a.N0max;
    p.row_mass = a.row_mass ? a.row_mass + (long long)b0 * a.N0max : nullptr;
    p.col_mass = a.col_mass ? a.col_mass + (long long)b0 * a.N1max : nullptr;
    p.n0s = a.n0s ? a.n0s + b0 : nullptr;
    p.n1s = a.n1s ? a.n1s + b0 : nullptr;
    p.N0max = a.N0max;
    p.N1max = a.N1max;
    p.write_scores = a.write_scores;
    const int rows_per_cta = sk_rows_per_cta(R, nb, 2 * num_sms(), SKR_CONSUMERS);
    p.rows_per_cta = rows_per_cta;
    int slots = (int)((SKR_SMEM_BUDGET - fixed) / row_bytes);
    if (slots > 64) slots = 64;
    if (slots > rows_per_cta) slots = rows_per_cta;
    // Slot s must always be drained by the same consumer warp (row r -> warp r % 4, slot r % slots): otherwise a fast
    // warp could run a whole lap ahead of a slow one and mis-read the phase parity of a barrier it has never seen.
    slots = slots / SKR_CONSUMERS * SKR_CONSUMERS;
    IMP_REQUIRE(slots >= SKR_CONSUMERS, "sinkhorn: a row of %d floats does not fit the shared-memory ring", a.ldp);
    p.ring_slots = slots;
    const size_t smem = (size_t)slots * row_bytes + fixed;
    dim3 grid((R + rows_per_cta - 1) / rows_per_cta, nb);
    float* col[3] = {a.colbuf + (size_t)b0 * a.ldp, a.colbuf + ((size_t)a.batch + b0) * a.ldp,
                     a.colbuf + (2 * (size_t)a.batch + b0) * a.ldp};
    p.col_prev = nullptr;
    p.col_acc = col[0];
    p.col_zero = col[1];
    p.do_iter = iters > 0 ? 1 : 0;
    sk_ring_kernel<NV, SK_INIT><<<grid, SKR_THREADS, smem, st>>>(p);
    const bool prof = g_profile && b0 == 0 && nb == a.batch && iters > 1;
    if (prof) sk_profile_begin(st);
    for (int k = 1; k < iters; ++k) {
      p.col_prev = col[(k - 1) % 3];
      p.col_acc = col[k % 3];
      p.col_zero = col[(k + 1) % 3];
      sk_ring_kernel<NV, SK_ITER><<<grid, SKR_THREADS, smem, st>>>(p);
    }
    if (prof) sk_profile_end(st, iters - 1);
    const float* col_last = col[(iters > 0 ? iters - 1 : 0) % 3];
    p.col_prev = col_last;
    p.col_acc = nullptr;
    p.col_zero = nullptr;
    sk_ring_kernel<NV, SK_FINAL><<<grid, SKR_THREADS, smem, st>>>(p);
    const int slab = 256;
    sk_colmax_kernel<<<dim3((a.N1max + 127) / 128, (a.N0max + slab - 1) / slab, nb), 128, 0, st>>>(
        p.P, a.p_batch_stride, a.ldp, p.u, col_last, a.write_scores, iters > 0 ? 1 : 0,
        reinterpret_cast<unsigned long long*>(a.col_key) + (size_t)b0 * a.N1max, p.n0s, p.n1s, a.N0max, a.N1max, slab);
  }
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_sinkhorn(const SinkhornArgs& a, cudaStream_t st) {
  IMP_REQUIRE(a.batch > 0 && a.N0max > 0 && a.N1max > 0, "sinkhorn: empty problem");
  IMP_REQUIRE(a.ldp % 4 == 0 && a.ldp >= a.N1max + 1, "sinkhorn: ldp must be a multiple of 4 and >= N1+1");
  IMP_REQUIRE(a.iters >= 0, "sinkhorn: negative iteration count");
  IMP_REQUIRE(a.ldd % 4 == 0 && a.ldd >= ((a.N1max + 3) & ~3) && a.dist_batch_stride % 4 == 0 &&
                  (reinterpret_cast<uintptr_t>(a.dist) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.P) & 15) == 0,
              "sinkhorn: dist / P must be 16-byte aligned with row strides that are multiples of 4 floats "
              "(ldd %d, N1 %d)", a.ldd, a.N1max);
  const int C = a.N1max + 1;
  if (C <= 4 * 32 * 5) return run_sinkhorn<5>(a, st);
  if (C <= 4 * 32 * 9) return run_sinkhorn<9>(a, st);
  if (C <= 4 * 32 * 17) return run_sinkhorn<17>(a, st);
  if (C <= 4 * 32 * 25) return run_sinkhorn<25>(a, st);
  if (C <= 4096 && a.q_store != nullptr && a.row_stats != nullptr) {  // beyond the row-ring kernels: column-split path only
    IMP_CUDA_OK(cudaMemsetAsync(a.col_key, 0, (size_t)a.batch * a.N1max * sizeof(unsigned long long), st));
    if (a.col_mass) IMP_CUDA_OK(cudaMemsetAsync(a.col_mass, 0, (size_t)a.batch * a.N1max * sizeof(float), st));
    IMP_CUDA_OK(cudaMemsetAsync(a.colbuf, 0, (size_t)a.batch * a.ldp * sizeof(float), st));
    return run_sinkhorn_compact(a, st);
  }
  set_error("sinkhorn: N1 = %d exceeds the supported maximum of %d columns (%d with a q_store workspace)", a.N1max,
            4 * 32 * 25 - 1, 4095);
  return 2;
}

int launch_matches(const MatchArgs& m, cudaStream_t st) {
  const int n = m.N0max > m.N1max ? m.N0max : m.N1max;
  dim3 grid((n + 255) / 256, m.batch);
  sk_matches_kernel<<<grid, 256, 0, st>>>(m.row_max, m.row_arg, reinterpret_cast<const unsigned long long*>(m.col_key), m.p_thresh,
                                          reinterpret_cast<long long*>(m.indices0), reinterpret_cast<long long*>(m.indices1),
                                          m.mscores0, m.mscores1, m.n0s, m.n1s, m.N0max, m.N1max, m.out0_batch_stride,
                                          m.out1_batch_stride);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace imp
