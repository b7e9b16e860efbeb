// Pieces shared by the fp32 Sinkhorn kernels (sinkhorn.cu) and the compact-storage ones (sinkhorn_q.cu).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "ptx.cuh"
#include "sinkhorn.cuh"

namespace imp {

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

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float4 v_from_colsum(const float* __restrict__ colsum, int c0, int C) {
  // v_j = c_j / (colsum_j + eps);  c_j = 1, last real column C-1 has mass C;  pad columns -> 0
  const float4 s = *reinterpret_cast<const float4*>(colsum + c0);
  float4 v;
  v.x = (c0 + 0 < C) ? ((c0 + 0 == C - 1) ? (float)C : 1.f) / (s.x + SK_EPS) : 0.f;
  v.y = (c0 + 1 < C) ? ((c0 + 1 == C - 1) ? (float)C : 1.f) / (s.y + SK_EPS) : 0.f;
  v.z = (c0 + 2 < C) ? ((c0 + 2 == C - 1) ? (float)C : 1.f) / (s.z + SK_EPS) : 0.f;
  v.w = (c0 + 3 < C) ? ((c0 + 3 == C - 1) ? (float)C : 1.f) / (s.w + SK_EPS) : 0.f;
  return v;
}

// exp(x) for x <= 0 with ~3e-7 relative error in 6 instructions: x*log2(e) is split into a rounded product t and its
// exact residual e (FMA), 2^t comes from the MUFU and the residual is applied to first order.  (expf() costs ~20
// instructions and made the softmax pass compute-bound: 1.28 ms per launch instead of the ~0.33 ms its traffic needs.)
// Written with explicit rounding intrinsics so every kernel that re-derives a probability from its logit (init, final,
// column arg-max of the compact-storage path) produces the same bits.
__device__ __forceinline__ float sk_exp(float x) {
  x = fmaxf(x, -200.f);  // exp(-200) underflows to 0 in fp32; also keeps the residual finite for the -FLT_MAX pad
  const float t = __fmul_rn(x, 1.4426950408889634f);
  const float e = __fmaf_rn(x, 1.925963033500235e-8f, __fmaf_rn(x, 1.4426950408889634f, -t));
  return __fmul_rn(fast_exp2(t), __fmaf_rn(e, 0.6931471805599453f, 1.0f));
}

__device__ __forceinline__ unsigned long long pack_max_key(float val, int idx) {
  // scores are >= 0, so the raw bits order like the values; ~idx makes the LOWEST index win ties
  return (static_cast<unsigned long long>(__float_as_uint(val)) << 32) | (0xFFFFFFFFu - (unsigned)idx);
}

// wave-aware block height shared by both streaming paths: rows per CTA (multiple of `step`, 8..128) whose CTA count
// fills whole waves of `wave_ctas` resident CTAs best -- e.g. 64 x 2001 rows: 128-row blocks give 1024 CTAs = 3.46 waves, 88-row blocks
// 1472 CTAs = 4.97 waves.  Larger blocks win ties (fewer column-sum flushes).
inline int sk_rows_per_cta(int R, int nb, int wave_ctas, int step) {
  const long long wave = wave_ctas;
  int rows_per_cta = 8;
  double best = -1.0;
  for (int r = 128; r >= 8; r -= step) {
    const long long ctas = (long long)nb * ((R + r - 1) / r);
    const long long waves = (ctas + wave - 1) / wave;
    const double eff = (double)ctas / (double)(waves * wave);
    const double balance = (double)R / (double)(((R + r - 1) / r) * r);  // ragged last block of each matrix
    // every work item pays a fixed price next to its r rows (v of its columns recomputed from the column-sum buffers, the
    // column sums flushed with one atomic per column): weighted as 4 rows' worth, which keeps the measured optimum (88 rows)
    // for batch 64 x 2001.  Without this term batch 128 x 2001 chose 8-row items (32128 of them fill 109 waves to 99.6 %)
    // and ran 4x slower per matrix than batch 64 (measured: 0.690 vs 0.167 ms per sweep).
    const double overhead = (double)r / (double)(r + 4);
    const double score = eff * balance * overhead;
    if (score > best + 1e-3) {
      best = score;
      rows_per_cta = r;
    }
  }
  return rows_per_cta;
}

// bench.py's in-library timing of the iteration launches (see imp_set_profiling)
bool sk_profiling_on();
void sk_profile_begin(cudaStream_t st);
void sk_profile_end(cudaStream_t st, int launches);

int run_sinkhorn_compact(const SinkhornArgs& a, cudaStream_t st);  // sinkhorn_q.cu

}  // namespace imp
