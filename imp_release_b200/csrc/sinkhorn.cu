// Sinkhorn optimal-transport assignment, probability domain, bit-for-bit the reference's recurrence
// (nets/layers.py:27-46): pad a dustbin column + row with bin_score, p = softmax_rows(M), u = v = 1,
// S x { u = r / (sum_j p v + 1e-8);  v = c / (sum_i p u + 1e-8) },  out = (p u) v,
// r = [1..1, N0+1], c = [1..1, N1+1].  Row softmax is the max-subtracted (LSE-stabilised) form.
//
// HBM/L2-bound streaming kernels.  One launch per Sinkhorn iteration reads the matrix ONCE: a warp
// keeps a whole row in registers, reduces it against v (warp shuffles) to get u_i, then immediately
// folds p_ij * u_i into per-lane column accumulators; the CTA combines its warps in shared memory and
// issues one global atomicAdd per column.  v is never materialised: the next launch recomputes
// c_j / (colsum_j + eps) on the fly from the accumulated column sums (three rotating buffers: read /
// accumulate / being-zeroed).  The final launch applies (p u) v in place and fuses the row / column
// arg-max (lowest index wins ties, like torch.max on CPU) and the row / column masses EIMP's pooling needs.
#include "sinkhorn.cuh"

#include <float.h>

#include "common.h"
#include "ptx.cuh"

namespace imp {

static constexpr int SK_THREADS = 256;
static constexpr int SK_WARPS = SK_THREADS / 32;
static constexpr float SK_EPS = 1e-8f;

struct SkDims {
  int R, C;  // augmented rows / cols of this sample
};

__device__ __forceinline__ SkDims sk_dims(const int* n0s, const int* n1s, int b, int N0max, int N1max) {
  SkDims d;
  d.R = (n0s ? n0s[b] : N0max) + 1;
  d.C = (n1s ? n1s[b] : N1max) + 1;
  return d;
}

// Row loader: lane l owns float4 groups g = l + 32*k (columns 4g..4g+3), k < NV.
template <int NV>
__device__ __forceinline__ void load_row(const float* __restrict__ row, int C, float4 (&x)[NV]) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c0 = 4 * (lane_id() + 32 * k);
    if (c0 < C)  // ld is a multiple of 4 and the pad columns hold zeros, so a full float4 is always readable
      x[k] = *reinterpret_cast<const float4*>(row + c0);
    else
      x[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__device__ __forceinline__ float4 v_from_colsum(const float* __restrict__ colsum, int c0, int C) {
  // v_j = c_j / (colsum_j + eps);  c_j = 1, last real column C-1 has mass C;  pad columns -> 0
  float4 s = *reinterpret_cast<const float4*>(colsum + c0);
  float4 v;
  v.x = (c0 + 0 < C) ? ((c0 + 0 == C - 1) ? (float)C : 1.f) / (s.x + SK_EPS) : 0.f;
  v.y = (c0 + 1 < C) ? ((c0 + 1 == C - 1) ? (float)C : 1.f) / (s.y + SK_EPS) : 0.f;
  v.z = (c0 + 2 < C) ? ((c0 + 2 == C - 1) ? (float)C : 1.f) / (s.z + SK_EPS) : 0.f;
  v.w = (c0 + 3 < C) ? ((c0 + 3 == C - 1) ? (float)C : 1.f) / (s.w + SK_EPS) : 0.f;
  return v;
}

template <int NV>
__device__ __forceinline__ void flush_colacc(float4 (&acc)[NV], float* s_col, float* __restrict__ g_col, int C) {
  // combine the CTA's warps in shared memory, then one global atomic per column
  for (int j = threadIdx.x; j < ((C + 3) & ~3); j += SK_THREADS) s_col[j] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c0 = 4 * (lane_id() + 32 * k);
    if (c0 < C) {
      atomicAdd(s_col + c0 + 0, acc[k].x);
      atomicAdd(s_col + c0 + 1, acc[k].y);
      atomicAdd(s_col + c0 + 2, acc[k].z);
      atomicAdd(s_col + c0 + 3, acc[k].w);
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < C; j += SK_THREADS) atomicAdd(g_col + j, s_col[j]);
}

// ---------------------------------------------------------------------------------------------
// init: P = softmax_rows(pad(dist)); optionally also the first Sinkhorn half-steps (u with v = 1,
// column sums with that u).  Zeroes the NEXT column-sum buffer.
template <int NV>
__global__ void __launch_bounds__(SK_THREADS)
sk_init_kernel(const float* __restrict__ dist, long long dist_bs, int ldd, const float* __restrict__ bin_score,
               float* __restrict__ P, long long p_bs, int ldp, float* __restrict__ u, float* __restrict__ col_acc,
               float* __restrict__ col_zero, const int* __restrict__ n0s, const int* __restrict__ n1s, int N0max,
               int N1max, int rows_per_cta, int do_iter) {
  extern __shared__ float s_col[];
  const int b = blockIdx.y;
  const SkDims d = sk_dims(n0s, n1s, b, N0max, N1max);
  const int row0 = blockIdx.x * rows_per_cta;
  if (row0 >= d.R) return;
  const int warp = threadIdx.x >> 5;
  const float bin = *bin_score;
  float4 acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (blockIdx.x == 0 && col_zero != nullptr)
    for (int j = threadIdx.x; j < ldp; j += SK_THREADS) col_zero[(long long)b * ldp + j] = 0.f;

  const int row_end = min(row0 + rows_per_cta, d.R);
  for (int i = row0 + warp; i < row_end; i += SK_WARPS) {
    float4 x[NV];
    const bool bin_row = (i == d.R - 1);
    const float* drow = dist + b * dist_bs + (long long)i * ldd;
    float m = -FLT_MAX;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = 4 * (lane_id() + 32 * k);
      float e[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int c = c0 + t;
        float val = -FLT_MAX;
        if (c < d.C) val = (bin_row || c == d.C - 1) ? bin : drow[c];
        e[t] = val;
        m = fmaxf(m, val);
      }
      x[k] = make_float4(e[0], e[1], e[2], e[3]);
    }
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = 4 * (lane_id() + 32 * k);
      x[k].x = (c0 + 0 < d.C) ? expf(x[k].x - m) : 0.f;
      x[k].y = (c0 + 1 < d.C) ? expf(x[k].y - m) : 0.f;
      x[k].z = (c0 + 2 < d.C) ? expf(x[k].z - m) : 0.f;
      x[k].w = (c0 + 3 < d.C) ? expf(x[k].w - m) : 0.f;
      s += (x[k].x + x[k].y) + (x[k].z + x[k].w);
    }
    s = warp_sum(s);
    float* prow = P + b * p_bs + (long long)i * ldp;
    float rs = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = 4 * (lane_id() + 32 * k);
      x[k].x = x[k].x / s;
      x[k].y = x[k].y / s;
      x[k].z = x[k].z / s;
      x[k].w = x[k].w / s;
      rs += (x[k].x + x[k].y) + (x[k].z + x[k].w);
      if (c0 < ldp) *reinterpret_cast<float4*>(prow + c0) = x[k];
    }
    if (do_iter) {
      rs = warp_sum(rs);  // sum_j p_ij * v_j with v = 1
      const float ui = (bin_row ? (float)d.R : 1.f) / (rs + SK_EPS);
      if (lane_id() == 0) u[(long long)b * (N0max + 1) + i] = ui;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        acc[k].x += x[k].x * ui;
        acc[k].y += x[k].y * ui;
        acc[k].z += x[k].z * ui;
        acc[k].w += x[k].w * ui;
      }
    }
  }
  if (do_iter) flush_colacc<NV>(acc, s_col, col_acc + (long long)b * ldp, d.C);
}

// one full Sinkhorn iteration (u then column sums) in a single sweep over P
template <int NV>
__global__ void __launch_bounds__(SK_THREADS)
sk_iter_kernel(const float* __restrict__ P, long long p_bs, int ldp, const float* __restrict__ col_prev,
               float* __restrict__ col_acc, float* __restrict__ col_zero, float* __restrict__ u,
               const int* __restrict__ n0s, const int* __restrict__ n1s, int N0max, int N1max, int rows_per_cta) {
  extern __shared__ float s_col[];
  const int b = blockIdx.y;
  const SkDims d = sk_dims(n0s, n1s, b, N0max, N1max);
  const int row0 = blockIdx.x * rows_per_cta;
  if (row0 >= d.R) return;
  const int warp = threadIdx.x >> 5;
  if (blockIdx.x == 0)
    for (int j = threadIdx.x; j < ldp; j += SK_THREADS) col_zero[(long long)b * ldp + j] = 0.f;

  float4 acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  float* s_v = s_col + ldp;  // v_j for this sample, shared by the CTA's warps
  for (int c0 = 4 * threadIdx.x; c0 < ((d.C + 3) & ~3); c0 += 4 * SK_THREADS)
    *reinterpret_cast<float4*>(s_v + c0) = v_from_colsum(col_prev + (long long)b * ldp, c0, d.C);
  __syncthreads();
  const int row_end = min(row0 + rows_per_cta, d.R);
  for (int i = row0 + warp; i < row_end; i += SK_WARPS) {
    float4 x[NV];
    load_row<NV>(P + b * p_bs + (long long)i * ldp, d.C, x);
    float rs = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = 4 * (lane_id() + 32 * k);
      if (c0 < d.C) {
        const float4 v = *reinterpret_cast<const float4*>(s_v + c0);
        rs += (x[k].x * v.x + x[k].y * v.y) + (x[k].z * v.z + x[k].w * v.w);
      }
    }
    rs = warp_sum(rs);
    const float ui = ((i == d.R - 1) ? (float)d.R : 1.f) / (rs + SK_EPS);
    if (lane_id() == 0) u[(long long)b * (N0max + 1) + i] = ui;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      acc[k].x += x[k].x * ui;
      acc[k].y += x[k].y * ui;
      acc[k].z += x[k].z * ui;
      acc[k].w += x[k].w * ui;
    }
  }
  flush_colacc<NV>(acc, s_col, col_acc + (long long)b * ldp, d.C);
}

__device__ __forceinline__ unsigned long long pack_max_key(float val, int idx) {
  // scores are >= 0, so the raw bits order like the values; ~idx makes the LOWEST index win ties
  return (static_cast<unsigned long long>(__float_as_uint(val)) << 32) | (0xFFFFFFFFu - (unsigned)idx);
}

// final scaling out = (p u) v in place + fused row/col arg-max and masses over the non-dustbin block
template <int NV>
__global__ void __launch_bounds__(SK_THREADS)
sk_final_kernel(float* __restrict__ P, long long p_bs, int ldp, const float* __restrict__ col_last,
                const float* __restrict__ u, int has_iter, float* __restrict__ row_max, int* __restrict__ row_arg,
                unsigned long long* __restrict__ col_key, float* __restrict__ row_mass, float* __restrict__ col_mass,
                const int* __restrict__ n0s, const int* __restrict__ n1s, int N0max, int N1max, int rows_per_cta) {
  extern __shared__ float s_col[];  // [ldp] masses, then [ldp] u64 keys
  const int b = blockIdx.y;
  const SkDims d = sk_dims(n0s, n1s, b, N0max, N1max);
  const int row0 = blockIdx.x * rows_per_cta;
  if (row0 >= d.R) return;
  const int warp = threadIdx.x >> 5;
  const int ldp4 = (d.C + 3) & ~3;
  unsigned long long* s_key = reinterpret_cast<unsigned long long*>(s_col + ((ldp + 1) & ~1));
  for (int j = threadIdx.x; j < ldp4; j += SK_THREADS) {
    s_col[j] = 0.f;
    s_key[j] = 0ull;
  }

  float* s_v = reinterpret_cast<float*>(s_key + ldp);
  for (int c0 = 4 * threadIdx.x; c0 < ldp4; c0 += 4 * SK_THREADS)
    *reinterpret_cast<float4*>(s_v + c0) =
        has_iter ? v_from_colsum(col_last + (long long)b * ldp, c0, d.C) : make_float4(1.f, 1.f, 1.f, 1.f);
  __syncthreads();
  const int row_end = min(row0 + rows_per_cta, d.R);
  for (int i = row0 + warp; i < row_end; i += SK_WARPS) {
    float4 x[NV];
    float* prow = P + b * p_bs + (long long)i * ldp;
    load_row<NV>(prow, d.C, x);
    const float ui = has_iter ? u[(long long)b * (N0max + 1) + i] : 1.f;
    const bool inner_row = i < d.R - 1;
    float best = -1.f;
    int best_j = 0x7fffffff;
    float mass = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = 4 * (lane_id() + 32 * k);
      if (c0 >= d.C) continue;
      const float4 v = *reinterpret_cast<const float4*>(s_v + c0);
      float o[4] = {(x[k].x * ui) * v.x, (x[k].y * ui) * v.y, (x[k].z * ui) * v.z, (x[k].w * ui) * v.w};
      *reinterpret_cast<float4*>(prow + c0) = make_float4(o[0], o[1], o[2], o[3]);
      if (inner_row) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int c = c0 + t;
          if (c < d.C - 1) {
            mass += o[t];
            if (o[t] > best) {  // strict: keeps the lowest column among equal values in this lane
              best = o[t];
              best_j = c;
            }
            if (col_mass) atomicAdd(s_col + c, o[t]);
            atomicMax(s_key + c, pack_max_key(o[t], i));
          }
        }
      }
    }
    if (inner_row) {
      // warp arg-max with lowest-index tie break
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
        row_max[(long long)b * N0max + i] = best;
        row_arg[(long long)b * N0max + i] = best_j;
        if (row_mass) row_mass[(long long)b * N0max + i] = mass;
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < d.C - 1; j += SK_THREADS) {
    atomicMax(col_key + (long long)b * N1max + j, s_key[j]);
    if (col_mass) atomicAdd(col_mass + (long long)b * N1max + j, s_col[j]);
  }
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
  }
  if (t < n1 && indices1 != nullptr) {
    const int i = (int)(0xFFFFFFFFu - (unsigned)(ck[t] & 0xFFFFFFFFull));
    const bool mutual1 = ra[i] == t;  // then row i is mutual too and mscores0[i] = row_max[i]
    const float ms0_i = mutual1 ? rm[i] : 0.f;
    mscores1[b * out1_bs + t] = ms0_i;
    indices1[b * out1_bs + t] = (mutual1 && ms0_i > p_thresh) ? (long long)i : -1ll;
  }
}

template <int NV>
static int run_sinkhorn(const SinkhornArgs& a, cudaStream_t st) {
  const int R = a.N0max + 1;
  // aim for >= 2 waves of CTAs; a warp handles rows_per_cta / 8 rows
  int rows_per_cta = 32;
  while (rows_per_cta > 8 && (long long)a.batch * ((R + rows_per_cta - 1) / rows_per_cta) < 2LL * num_sms()) rows_per_cta >>= 1;
  dim3 grid((R + rows_per_cta - 1) / rows_per_cta, a.batch);
  const size_t smem_col = 2 * (size_t)a.ldp * sizeof(float);
  const size_t smem_fin = (size_t)((a.ldp + 1) & ~1) * sizeof(float) + (size_t)a.ldp * (sizeof(unsigned long long) + sizeof(float));
  float* col[3] = {a.colbuf, a.colbuf + (size_t)a.batch * a.ldp, a.colbuf + 2 * (size_t)a.batch * a.ldp};
  const int iters = a.iters;
  IMP_CUDA_OK(cudaMemsetAsync(col[0], 0, (size_t)a.batch * a.ldp * sizeof(float), st));
  sk_init_kernel<NV><<<grid, SK_THREADS, smem_col, st>>>(a.dist, a.dist_batch_stride, a.ldd, a.bin_score, a.P,
                                                          a.p_batch_stride, a.ldp, a.u, col[0], col[1], a.n0s, a.n1s,
                                                          a.N0max, a.N1max, rows_per_cta, iters > 0 ? 1 : 0);
  for (int k = 1; k < iters; ++k) {
    sk_iter_kernel<NV><<<grid, SK_THREADS, smem_col, st>>>(a.P, a.p_batch_stride, a.ldp, col[(k - 1) % 3], col[k % 3],
                                                            col[(k + 1) % 3], a.u, a.n0s, a.n1s, a.N0max, a.N1max,
                                                            rows_per_cta);
  }
  IMP_CUDA_OK(cudaMemsetAsync(a.col_key, 0, (size_t)a.batch * a.N1max * sizeof(unsigned long long), st));
  if (a.col_mass) IMP_CUDA_OK(cudaMemsetAsync(a.col_mass, 0, (size_t)a.batch * a.N1max * sizeof(float), st));
  sk_final_kernel<NV><<<grid, SK_THREADS, smem_fin, st>>>(a.P, a.p_batch_stride, a.ldp,
                                                           col[(iters > 0 ? iters - 1 : 0) % 3], a.u, iters > 0 ? 1 : 0,
                                                           a.row_max, a.row_arg, reinterpret_cast<unsigned long long*>(a.col_key), a.row_mass, a.col_mass,
                                                           a.n0s, a.n1s, a.N0max, a.N1max, rows_per_cta);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_sinkhorn(const SinkhornArgs& a, cudaStream_t st) {
  IMP_REQUIRE(a.batch > 0 && a.N0max > 0 && a.N1max > 0, "sinkhorn: empty problem");
  IMP_REQUIRE(a.ldp % 4 == 0 && a.ldp >= a.N1max + 1, "sinkhorn: ldp must be a multiple of 4 and >= N1+1");
  IMP_REQUIRE(a.iters >= 0, "sinkhorn: negative iteration count");
  const int C = a.N1max + 1;
  if (C <= 4 * 32 * 5) return run_sinkhorn<5>(a, st);
  if (C <= 4 * 32 * 9) return run_sinkhorn<9>(a, st);
  if (C <= 4 * 32 * 17) return run_sinkhorn<17>(a, st);
  if (C <= 4 * 32 * 33) return run_sinkhorn<33>(a, st);
  set_error("sinkhorn: N1 = %d exceeds the supported maximum of %d columns", a.N1max, 4 * 32 * 33 - 1);
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
