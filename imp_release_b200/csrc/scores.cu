// Secondary scoring paths of the boundary API:
//  * dual_softmax (nets/layers.py:20-24), selected by with_sinkhorn=False / --use_dual_softmax
//  * row / column arg-max of a caller-provided score matrix, feeding GM.compute_matches (nets/gm.py:305-320) when
//    it is invoked on an arbitrary tensor (eval/matching.py:119 re-thresholds the last scores with p=0.2).
#include <float.h>

#include "common.h"
#include "ptx.cuh"

namespace imp {

__device__ __forceinline__ unsigned long long pack_key(float val, int idx) {
  return (static_cast<unsigned long long>(__float_as_uint(val)) << 32) | (0xFFFFFFFFu - (unsigned)idx);
}

// rows: one warp per row over the non-dustbin block [N0, N1]
__global__ void row_argmax_kernel(const float* __restrict__ P, long long p_bs, int ldp, float* __restrict__ row_max,
                                  int* __restrict__ row_arg, float* __restrict__ row_mass, int N0max, int N1max,
                                  const int* __restrict__ n0s, const int* __restrict__ n1s) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int N0 = N0max;  // output stride
  const int n0 = n0s ? n0s[b] : N0max, N1 = n1s ? n1s[b] : N1max;
  if (i >= n0) return;
  const float* row = P + b * p_bs + (long long)i * ldp;
  float best = -FLT_MAX, mass = 0.f;
  int bj = 0x7fffffff;
  for (int j = lane_id(); j < N1; j += 32) {
    const float v = row[j];
    mass += v;
    if (v > best) {
      best = v;
      bj = j;
    }
  }
  mass = warp_sum(mass);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (ob > best || (ob == best && oj < bj)) {
      best = ob;
      bj = oj;
    }
  }
  if (lane_id() == 0) {
    row_max[(long long)b * N0 + i] = best;
    row_arg[(long long)b * N0 + i] = bj;
    if (row_mass) row_mass[(long long)b * N0 + i] = mass;
  }
}

// columns: thread per column over a slab of rows, merged with a packed atomicMax (lowest row wins ties)
__global__ void col_argmax_kernel(const float* __restrict__ P, long long p_bs, int ldp,
                                  unsigned long long* __restrict__ col_key, float* __restrict__ col_mass, int N0max,
                                  int N1max, const int* __restrict__ n0s, const int* __restrict__ n1s, int rows_per_slab) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int N1 = N1max;  // output stride
  if (j >= (n1s ? n1s[b] : N1max)) return;
  const int i0 = blockIdx.y * rows_per_slab;
  const int i1 = min(i0 + rows_per_slab, n0s ? n0s[b] : N0max);
  float best = -FLT_MAX, mass = 0.f;
  int bi = 0;
  for (int i = i0; i < i1; ++i) {
    const float v = P[b * p_bs + (long long)i * ldp + j];
    mass += v;
    if (v > best) {
      best = v;
      bi = i;
    }
  }
  if (i1 > i0) {
    atomicMax(col_key + (long long)b * N1 + j, pack_key(fmaxf(best, 0.f), bi));
    if (col_mass) atomicAdd(col_mass + (long long)b * N1 + j, mass);
  }
}

int launch_score_argmax(const float* P, long long p_bs, int ldp, float* row_max, int* row_arg,
                        unsigned long long* col_key, float* row_mass, float* col_mass, int N0, int N1, int batch,
                        const int* n0s, const int* n1s, cudaStream_t st) {
  if (batch == 0 || N0 == 0 || N1 == 0) return 0;
  IMP_CUDA_OK(cudaMemsetAsync(col_key, 0, (size_t)batch * N1 * sizeof(unsigned long long), st));
  if (col_mass) IMP_CUDA_OK(cudaMemsetAsync(col_mass, 0, (size_t)batch * N1 * sizeof(float), st));
  row_argmax_kernel<<<dim3((N0 + 7) / 8, batch), 256, 0, st>>>(P, p_bs, ldp, row_max, row_arg, row_mass, N0, N1, n0s, n1s);
  const int slab = 64;
  col_argmax_kernel<<<dim3((N1 + 127) / 128, (N0 + slab - 1) / slab, batch), 128, 0, st>>>(P, p_bs, ldp, col_key,
                                                                                           col_mass, N0, N1, n0s, n1s, slab);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------- dual softmax
__device__ __forceinline__ float padded_at(const float* __restrict__ dist, int ldd, float bin, int i, int j, int N0,
                                           int N1) {
  return (i == N0 || j == N1) ? bin : dist[(long long)i * ldd + j];
}

__global__ void ds_row_lse_kernel(const float* __restrict__ dist, long long d_bs, int ldd,
                                  const float* __restrict__ bin_score, float* __restrict__ row_lse, int N0max, int N1max,
                                  const int* __restrict__ n0s, const int* __restrict__ n1s) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int N0 = n0s ? n0s[b] : N0max, N1 = n1s ? n1s[b] : N1max;
  if (i > N0) return;
  const float bin = *bin_score;
  const float* d = dist + b * d_bs;
  float m = -FLT_MAX;
  for (int j = lane_id(); j <= N1; j += 32) m = fmaxf(m, padded_at(d, ldd, bin, i, j, N0, N1));
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane_id(); j <= N1; j += 32) s += expf(padded_at(d, ldd, bin, i, j, N0, N1) - m);
  s = warp_sum(s);
  if (lane_id() == 0) row_lse[(long long)b * (N0max + 1) + i] = m + logf(s);
}

__global__ void ds_col_lse_kernel(const float* __restrict__ dist, long long d_bs, int ldd,
                                  const float* __restrict__ bin_score, float* __restrict__ col_lse, int N0max, int N1max,
                                  const int* __restrict__ n0s, const int* __restrict__ n1s) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int N0 = n0s ? n0s[b] : N0max, N1 = n1s ? n1s[b] : N1max;
  if (j > N1) return;
  const float bin = *bin_score;
  const float* d = dist + b * d_bs;
  float m = -FLT_MAX, s = 0.f;
  for (int i = 0; i <= N0; ++i) {
    const float v = padded_at(d, ldd, bin, i, j, N0, N1);
    if (v > m) {
      s = s * expf(m - v) + 1.f;
      m = v;
    } else {
      s += expf(v - m);
    }
  }
  col_lse[(long long)b * (N1max + 1) + j] = m + logf(s);
}

__global__ void ds_apply_kernel(const float* __restrict__ dist, long long d_bs, int ldd,
                                const float* __restrict__ bin_score, const float* __restrict__ row_lse,
                                const float* __restrict__ col_lse, float* __restrict__ P, long long p_bs, int ldp, int N0max,
                                int N1max, const int* __restrict__ n0s, const int* __restrict__ n1s) {
  const int b = blockIdx.z;
  const int i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ldp) return;
  const int N0 = n0s ? n0s[b] : N0max, N1 = n1s ? n1s[b] : N1max;
  float out = 0.f;
  if (i <= N0 && j <= N1) {
    const float x = padded_at(dist + b * d_bs, ldd, *bin_score, i, j, N0, N1);
    // exp(log_softmax_row + log_softmax_col)
    out = expf((x - row_lse[(long long)b * (N0max + 1) + i]) + (x - col_lse[(long long)b * (N1max + 1) + j]));
  }
  P[b * p_bs + (long long)i * ldp + j] = out;
}

// per-sample sizes n0s / n1s (may be NULL): sample b scores its leading n0s[b] x n1s[b] block (+ dustbins right behind
// it, like a stand-alone call on that block would); everything outside is written as 0
int launch_dual_softmax(const float* dist, long long d_bs, int ldd, const float* bin_score, float* P, long long p_bs,
                        int ldp, float* row_lse, float* col_lse, int N0, int N1, int batch, const int* n0s, const int* n1s,
                        cudaStream_t st) {
  if (batch == 0) return 0;
  ds_row_lse_kernel<<<dim3((N0 + 1 + 7) / 8, batch), 256, 0, st>>>(dist, d_bs, ldd, bin_score, row_lse, N0, N1, n0s, n1s);
  ds_col_lse_kernel<<<dim3((N1 + 1 + 127) / 128, batch), 128, 0, st>>>(dist, d_bs, ldd, bin_score, col_lse, N0, N1, n0s, n1s);
  ds_apply_kernel<<<dim3((ldp + 255) / 256, N0 + 1, batch), 256, 0, st>>>(dist, d_bs, ldd, bin_score, row_lse, col_lse, P,
                                                                          p_bs, ldp, N0, N1, n0s, n1s);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace imp
