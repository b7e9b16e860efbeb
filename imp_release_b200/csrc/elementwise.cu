// Streaming (HBM-bound) helpers around the tensor-core kernels: fp32 -> fp16 hi/lo plane split,
// instance-norm statistics + normalise/ReLU/split, keypoint-encoder small layers, row gathers.
#include "elementwise.cuh"

#include "common.h"
#include "ptx.cuh"

#include <stdlib.h>

namespace imp {

// x[n] (fp32) (+ optional addend y[n]) -> hi/lo fp16 planes.  4 elements per thread, 16 B loads.
__global__ void split_planes_kernel(const float* __restrict__ x, const float* __restrict__ y, __half* __restrict__ hi,
                                    __half* __restrict__ lo, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<const float4*>(x)[i];
  if (y != nullptr) {
    const float4 w = reinterpret_cast<const float4*>(y)[i];
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  __half h[4], l[4];
  split_f16x2(v.x, h[0], l[0]);
  split_f16x2(v.y, h[1], l[1]);
  split_f16x2(v.z, h[2], l[2]);
  split_f16x2(v.w, h[3], l[3]);
  reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
  reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
}

int launch_split_planes(const float* x, const float* addend, void* hi, void* lo, long long n, cudaStream_t st) {
  IMP_REQUIRE(n % 4 == 0, "split_planes: element count must be a multiple of 4");
  const long long n4 = n / 4;
  if (n4 == 0) return 0;
  split_planes_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(x, addend, reinterpret_cast<__half*>(hi),
                                                                    reinterpret_cast<__half*>(lo), n4);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

// hi/lo planes -> fp32 (debug / boundary export)
__global__ void merge_planes_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, float* __restrict__ x,
                                    long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = __half2float(hi[i]) + __half2float(lo[i]);
}

int launch_merge_planes(const void* hi, const void* lo, float* x, long long n, cudaStream_t st) {
  if (n == 0) return 0;
  merge_planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const __half*>(hi),
                                                                  reinterpret_cast<const __half*>(lo), x, n);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Instance norm over the token axis (InstanceNorm1d(eps=1e-3, affine=False), nets/layers.py:68) + ReLU,
// written as fp16 hi/lo planes for the next GEMM.  H: [B, N, C] fp32 token-major.
// One CTA owns 32 channels of one sample: pass 1 accumulates shifted sums (shift = first token's value,
// which removes the E[x^2]-E[x]^2 cancellation), pass 2 re-reads the slab (L2-resident) and normalises.
static constexpr int IN_CH = 32;
static constexpr int IN_THREADS = 256;

__global__ void __launch_bounds__(IN_THREADS)
instnorm_relu_split_kernel(const float* __restrict__ H, long long h_bs, int ldh, const int* __restrict__ ns, int Nmax,
                           int C, float eps, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                           float* __restrict__ out_f32, long long o_bs, int ldo, int relu) {
  __shared__ float s_sum[IN_THREADS / 32][IN_CH], s_sq[IN_THREADS / 32][IN_CH];
  __shared__ float s_mean[IN_CH], s_rstd[IN_CH];
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * IN_CH;
  const int n = ns ? ns[b] : Nmax;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = c0 + lane;
  const bool cok = c < C;
  const float* h = H + b * h_bs + c;
  const float shift = (cok && n > 0) ? h[0] : 0.f;
  float s = 0.f, q = 0.f;
  for (int t = warp; t < n; t += IN_THREADS / 32) {
    const float v = cok ? h[(long long)t * ldh] - shift : 0.f;
    s += v;
    q += v * v;
  }
  s_sum[warp][lane] = s;
  s_sq[warp][lane] = q;
  __syncthreads();
  if (warp == 0) {
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int w = 0; w < IN_THREADS / 32; ++w) {
      ts += s_sum[w][lane];
      tq += s_sq[w][lane];
    }
    const float inv_n = n > 0 ? 1.f / (float)n : 0.f;
    const float m = ts * inv_n;
    const float var = fmaxf(tq * inv_n - m * m, 0.f);  // biased variance
    s_mean[lane] = m + shift;
    s_rstd[lane] = 1.f / sqrtf(var + eps);
  }
  __syncthreads();
  const float mean = s_mean[lane], rstd = s_rstd[lane];
  for (int t = warp; t < n; t += IN_THREADS / 32) {
    if (!cok) continue;
    float v = (h[(long long)t * ldh] - mean) * rstd;
    if (relu) v = fmaxf(v, 0.f);
    const long long o = b * o_bs + (long long)t * ldo + c;
    if (out_f32 != nullptr) {
      out_f32[o] = v;
    } else {
      __half a, d;
      split_f16x2(v, a, d);
      out_hi[o] = a;
      out_lo[o] = d;
    }
  }
}

// Slab variant for the wide MLP hidden layer: one CTA owns 16 channels of one image and keeps the whole [N, 16] slab
// (128 KB at N = 2000) in shared memory, so H is read from HBM exactly once (stats and normalisation both run out of
// smem) and the hi/lo planes are written once: 12 B per element, the minimum for a stand-alone pass.
static constexpr int INS_THREADS = 512;

template <int INS_CH, int MINB>
__global__ void __launch_bounds__(INS_THREADS, MINB)
instnorm_slab_kernel(const float* __restrict__ H, long long h_bs, int ldh, const int* __restrict__ ns, int Nmax, float eps,
                     __half* __restrict__ out_hi, __half* __restrict__ out_lo, long long o_bs, int ldo, int relu) {
  extern __shared__ __align__(16) float slab[];  // [n][16]
  __shared__ float s_part[INS_THREADS / 32][2][INS_CH];
  __shared__ float s_mean[INS_CH], s_rstd[INS_CH];
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * INS_CH;
  const int n = ns ? ns[b] : Nmax;
  constexpr int QPT = INS_CH / 4;           // float4 groups per token
  constexpr int QSH = (QPT == 4) ? 2 : 1;   // log2(QPT)
  const int q = threadIdx.x & (QPT - 1);    // which float4 of the slab's channels this thread always handles
  const float* h = H + b * h_bs + c0;
  const float4 shift = n > 0 ? *reinterpret_cast<const float4*>(h + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), sq = s;
  const int total = n * QPT;
#pragma unroll 8
  for (int i = threadIdx.x; i < total; i += INS_THREADS) {
    const int t = i >> QSH;
    const float4 v = *reinterpret_cast<const float4*>(h + (long long)t * ldh + 4 * q);
    *reinterpret_cast<float4*>(slab + t * INS_CH + 4 * q) = v;
    const float4 d = make_float4(v.x - shift.x, v.y - shift.y, v.z - shift.z, v.w - shift.w);
    s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
    sq.x += d.x * d.x; sq.y += d.y * d.y; sq.z += d.z * d.z; sq.w += d.w * d.w;
  }
  // lanes with equal q: xor QPT, 2 QPT, ...
#pragma unroll
  for (int o = QPT; o < 32; o <<= 1) {
    s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
    s.z += __shfl_xor_sync(0xffffffffu, s.z, o); s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
    sq.x += __shfl_xor_sync(0xffffffffu, sq.x, o); sq.y += __shfl_xor_sync(0xffffffffu, sq.y, o);
    sq.z += __shfl_xor_sync(0xffffffffu, sq.z, o); sq.w += __shfl_xor_sync(0xffffffffu, sq.w, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < QPT) {
    float* ps = &s_part[warp][0][4 * lane];
    float* pq = &s_part[warp][1][4 * lane];
    ps[0] = s.x; ps[1] = s.y; ps[2] = s.z; ps[3] = s.w;
    pq[0] = sq.x; pq[1] = sq.y; pq[2] = sq.z; pq[3] = sq.w;
  }
  __syncthreads();
  if (threadIdx.x < INS_CH) {
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int w = 0; w < INS_THREADS / 32; ++w) {
      ts += s_part[w][0][threadIdx.x];
      tq += s_part[w][1][threadIdx.x];
    }
    const float inv_n = n > 0 ? 1.f / (float)n : 0.f;
    const float m = ts * inv_n;
    const float var = fmaxf(tq * inv_n - m * m, 0.f);  // biased variance of the shifted data
    const float sh = n > 0 ? h[threadIdx.x] : 0.f;
    s_mean[threadIdx.x] = m + sh;
    s_rstd[threadIdx.x] = 1.f / sqrtf(var + eps);
  }
  __syncthreads();
  const float4 mean = *reinterpret_cast<const float4*>(&s_mean[4 * q]);
  const float4 rstd = *reinterpret_cast<const float4*>(&s_rstd[4 * q]);
  __half* oh = out_hi + b * o_bs + c0 + 4 * q;
  __half* ol = out_lo + b * o_bs + c0 + 4 * q;
#pragma unroll 4
  for (int i = threadIdx.x; i < total; i += INS_THREADS) {
    const int t = i >> QSH;
    const float4 v = *reinterpret_cast<const float4*>(slab + t * INS_CH + 4 * q);
    float y[4] = {(v.x - mean.x) * rstd.x, (v.y - mean.y) * rstd.y, (v.z - mean.z) * rstd.z, (v.w - mean.w) * rstd.w};
    __half hh[4], ll[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (relu) y[k] = fmaxf(y[k], 0.f);
      split_f16x2(y[k], hh[k], ll[k]);
    }
    *reinterpret_cast<uint2*>(oh + (long long)t * ldo) = *reinterpret_cast<uint2*>(hh);
    *reinterpret_cast<uint2*>(ol + (long long)t * ldo) = *reinterpret_cast<uint2*>(ll);
  }
}

int launch_instnorm_relu_split(const float* H, long long h_bs, int ldh, const int* ns, int Nmax, int C, int batch,
                               float eps, int relu, void* out_hi, void* out_lo, float* out_f32, long long o_bs,
                               int ldo, cudaStream_t st) {
  if (batch == 0 || Nmax == 0) return 0;
  const bool aligned = out_f32 == nullptr && C % 16 == 0 && C >= 256 && ldh % 4 == 0 && ldo % 4 == 0 && h_bs % 4 == 0 &&
                       o_bs % 4 == 0 && (reinterpret_cast<uintptr_t>(H) & 15) == 0;
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("IMP_IN_VARIANT");
    variant = e ? atoi(e) : 1;
  }
  // Measured on B200 (tools/in_probe.py, 128 x 2000 x 512): 16-channel slabs, one CTA/SM: 0.315 ms (3.3 TB/s);
  // 8-channel slabs, two CTAs/SM: 0.523 ms -- 32-byte accesses per token waste half of every DRAM burst.  So the
  // 8-channel variant is opt-in (IMP_IN_VARIANT=0) only.
  if (aligned && variant == 0 && (size_t)Nmax * 8 * sizeof(float) <= 100 * 1024) {
    auto kern = instnorm_slab_kernel<8, 2>;
    static DeviceOnce configured;
    if (configured.first()) {
      IMP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      IMP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    kern<<<dim3(C / 8, batch), INS_THREADS, (size_t)Nmax * 8 * sizeof(float), st>>>(
        H, h_bs, ldh, ns, Nmax, eps, reinterpret_cast<__half*>(out_hi), reinterpret_cast<__half*>(out_lo), o_bs, ldo, relu);
    IMP_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (aligned && (size_t)Nmax * 16 * sizeof(float) <= 200 * 1024) {
    auto kern = instnorm_slab_kernel<16, 1>;
    static DeviceOnce configured;
    if (configured.first()) {
      IMP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    kern<<<dim3(C / 16, batch), INS_THREADS, (size_t)Nmax * 16 * sizeof(float), st>>>(
        H, h_bs, ldh, ns, Nmax, eps, reinterpret_cast<__half*>(out_hi), reinterpret_cast<__half*>(out_lo), o_bs, ldo, relu);
    IMP_CUDA_OK(cudaGetLastError());
    return 0;
  }
  dim3 grid((C + IN_CH - 1) / IN_CH, batch);
  instnorm_relu_split_kernel<<<grid, IN_THREADS, 0, st>>>(H, h_bs, ldh, ns, Nmax, C, eps,
                                                          reinterpret_cast<__half*>(out_hi),
                                                          reinterpret_cast<__half*>(out_lo), out_f32, o_bs, ldo,
                                                          relu);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Fused instance norm, second half (the first half -- per-tile column sums -- is the epilogue of the GEMM that produced H,
// gemm.cu).  finalize: one thread per (image, channel) adds the image's tile partials in tile order in fp64 (plain sums
// are fine at that precision: the relative error of the variance is ~1e-7 (1 + mean^2/var)) -> (mean, rstd).
__global__ void instnorm_finalize_kernel(const float2* __restrict__ partial, const float2* __restrict__ straddle,
                                         const int* __restrict__ ns, int Np, int C, float eps, float2* __restrict__ stats) {
  const int img = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int n = ns[img];
  const long long r0 = (long long)img * Np;
  const int t0 = (int)(r0 / 128), t1 = (int)((r0 + Np - 1) / 128);
  double s1 = 0.0, s2 = 0.0;
  for (int t = t0; t <= t1; ++t) {
    if ((long long)t * 128 >= r0) {  // a tile whose first image is this one
      const float2 v = partial[(long long)t * C + c];
      s1 += v.x;
      s2 += v.y;
    } else {                         // the tile that starts in the previous image and reaches into this one
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 v = straddle[((long long)img * 4 + q) * C + c];
        s1 += v.x;
        s2 += v.y;
      }
    }
  }
  const double inv_n = n > 0 ? 1.0 / n : 0.0;
  const double m = s1 * inv_n;
  const double var = fmax(s2 * inv_n - m * m, 0.0);  // biased variance
  stats[(long long)img * C + c] = make_float2((float)m, (float)(1.0 / sqrt(var + (double)eps)));
}

// apply: pure streaming pass, 8 channels per thread (2 x 16 B in, 2 x 16 B out); the (mean, rstd) pairs of a thread's fixed
// channel group stay in registers across its grid-stride rows
__global__ void __launch_bounds__(256)
instnorm_apply_kernel(const float* __restrict__ H, const float2* __restrict__ stats, int Np, int C, int relu,
                      __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
  const int img = blockIdx.y;
  const int cg = C >> 3;                        // channel groups of 8 per row
  const int rows_per_pass = blockDim.x / cg;    // host guarantees blockDim.x % cg == 0
  const int g = threadIdx.x % cg, r_in = threadIdx.x / cg;
  const float2* st = stats + (long long)img * C + g * 8;
  float mean[8], rstd[8];
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    const float4 m = *reinterpret_cast<const float4*>(st + k);  // (mean, rstd) x 2
    mean[k] = m.x; rstd[k] = m.y; mean[k + 1] = m.z; rstd[k + 1] = m.w;
  }
  for (int row = blockIdx.x * rows_per_pass + r_in; row < Np; row += gridDim.x * rows_per_pass) {
    const long long e = ((long long)img * Np + row) * C + g * 8;
    const float4 a = *reinterpret_cast<const float4*>(H + e);
    const float4 b = *reinterpret_cast<const float4*>(H + e + 4);
    float y[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    __half hh[8], ll[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      y[k] = (y[k] - mean[k]) * rstd[k];
      if (relu) y[k] = fmaxf(y[k], 0.f);
      split_f16x2(y[k], hh[k], ll[k]);
    }
    *reinterpret_cast<uint4*>(out_hi + e) = *reinterpret_cast<uint4*>(hh);
    *reinterpret_cast<uint4*>(out_lo + e) = *reinterpret_cast<uint4*>(ll);
  }
}

int launch_instnorm_apply(const float* H, const float* stat_partial, const float* stat_straddle, const int* ns, int Np, int C,
                          int images, float eps, int relu, float* stats, void* out_hi, void* out_lo, cudaStream_t st) {
  IMP_REQUIRE(Np >= 128 && images > 0, "instnorm_apply: needs images of >= 128 rows");
  instnorm_finalize_kernel<<<dim3((C + 127) / 128, images), 128, 0, st>>>(
      reinterpret_cast<const float2*>(stat_partial), reinterpret_cast<const float2*>(stat_straddle), ns, Np, C, eps,
      reinterpret_cast<float2*>(stats));
  if (out_hi == nullptr) {  // statistics only: the consumer GEMM normalises its A operand itself
    IMP_CUDA_OK(cudaGetLastError());
    return 0;
  }
  const int cg = C / 8;
  IMP_REQUIRE(C % 8 == 0 && cg <= 256 && 256 % cg == 0, "instnorm_apply: C / 8 must divide 256 (C = %d)", C);
  const int rows_per_pass = 256 / cg;
  int bx = (Np + rows_per_pass - 1) / rows_per_pass;
  const int cap = (num_sms() * 8 + images - 1) / images;  // ~8 CTAs per SM in total: grid-stride above that
  if (bx > cap) bx = cap < 1 ? 1 : cap;
  instnorm_apply_kernel<<<dim3(bx, images), 256, 0, st>>>(H, reinterpret_cast<const float2*>(stats), Np, C, relu,
                                                          reinterpret_cast<__half*>(out_hi), reinterpret_cast<__half*>(out_lo));
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Small fp32 linear layer for the keypoint encoder's narrow layers (Cin <= 64): y[t, o] = b[o] + sum_c x[t,c] w[o,c].
// The 3->32 and 32->64 layers are <2% of the encoder FLOPs; wider layers go through the tensor-core GEMM.
__global__ void small_linear_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                    const float* __restrict__ bias, float* __restrict__ Y, int ldy, long long rows,
                                    int Cin, int Cout) {
  extern __shared__ float s_w[];  // transposed [Cin][Cout] (lanes = consecutive outputs -> conflict-free) + [Cout]
  for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) {
    const int o = i / Cin, c = i - o * Cin;
    s_w[c * Cout + o] = W[i];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) s_w[Cout * Cin + i] = bias[i];
  __syncthreads();
  const long long total = rows * Cout;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx / Cout;
    const int o = (int)(idx - t * Cout);
    const float* x = X + t * ldx;
    float acc = s_w[Cout * Cin + o];
#pragma unroll 4
    for (int cidx = 0; cidx < Cin; ++cidx) acc = fmaf(__ldg(x + cidx), s_w[cidx * Cout + o], acc);
    Y[t * ldy + o] = acc;
  }
}

int launch_small_linear(const float* X, int ldx, const float* W, const float* bias, float* Y, int ldy, long long rows,
                        int Cin, int Cout, cudaStream_t st) {
  if (rows == 0) return 0;
  const size_t smem = (size_t)(Cout * Cin + Cout) * sizeof(float);
  IMP_REQUIRE(smem <= 48 * 1024, "small_linear: weight does not fit in shared memory");
  const long long total = rows * Cout;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  small_linear_kernel<<<blocks, 256, smem, st>>>(X, ldx, W, bias, Y, ldy, rows, Cin, Cout);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

// keypoint encoder input: [x_norm, y_norm, score] rows (nets/layers.py:88-90) from kpts [T,2] and scores [T]
__global__ void kenc_input_kernel(const float* __restrict__ kpts, const float* __restrict__ scores,
                                  float* __restrict__ out, long long T) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  out[4 * t + 0] = kpts[2 * t + 0];
  out[4 * t + 1] = kpts[2 * t + 1];
  out[4 * t + 2] = scores[t];
  out[4 * t + 3] = 0.f;
}

int launch_kenc_input(const float* kpts, const float* scores, float* out, long long T, cudaStream_t st) {
  if (T == 0) return 0;
  kenc_input_kernel<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(kpts, scores, out, T);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Row gather (EIMP compaction): out[b, r, :] = in[b, ids[b, r], :] for r < cnt[b]; 16-byte vectors.
__global__ void gather_rows_kernel(const uint4* __restrict__ in, long long in_bs, int ld_in,
                                   const int* __restrict__ ids, int ids_ld, const int* __restrict__ cnt,
                                   uint4* __restrict__ out, long long out_bs, int ld_out, int vec_per_row) {
  const int b = blockIdx.z;
  const int r = blockIdx.y;
  if (r >= cnt[b]) return;
  const int src = ids[(long long)b * ids_ld + r];
  for (int vi = threadIdx.x; vi < vec_per_row; vi += blockDim.x)
    out[b * out_bs + (long long)r * ld_out + vi] = in[b * in_bs + (long long)src * ld_in + vi];
}

int launch_gather_rows(const void* in, long long in_bs_bytes, int row_bytes_in, const int* ids, int ids_ld,
                       const int* cnt, void* out, long long out_bs_bytes, int row_bytes_out, int copy_bytes,
                       int max_rows, int batch, cudaStream_t st) {
  IMP_REQUIRE(copy_bytes % 16 == 0 && row_bytes_in % 16 == 0 && row_bytes_out % 16 == 0, "gather_rows: rows must be 16-byte multiples");
  if (batch == 0 || max_rows == 0) return 0;
  dim3 grid(1, max_rows, batch);
  gather_rows_kernel<<<grid, 64, 0, st>>>(reinterpret_cast<const uint4*>(in), in_bs_bytes / 16, row_bytes_in / 16, ids,
                                          ids_ld, cnt, reinterpret_cast<uint4*>(out), out_bs_bytes / 16,
                                          row_bytes_out / 16, copy_bytes / 16);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace imp
