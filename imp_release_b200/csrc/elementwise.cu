// Streaming (HBM-bound) helpers around the tensor-core kernels: fp32 -> fp16 hi/lo plane split,
// instance-norm statistics + normalise/ReLU/split, keypoint-encoder small layers, row gathers.
#include "elementwise.cuh"

#include "common.h"
#include "ptx.cuh"

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

int launch_instnorm_relu_split(const float* H, long long h_bs, int ldh, const int* ns, int Nmax, int C, int batch,
                               float eps, int relu, void* out_hi, void* out_lo, float* out_f32, long long o_bs,
                               int ldo, cudaStream_t st) {
  if (batch == 0 || Nmax == 0) return 0;
  dim3 grid((C + IN_CH - 1) / IN_CH, batch);
  instnorm_relu_split_kernel<<<grid, IN_THREADS, 0, st>>>(H, h_bs, ldh, ns, Nmax, C, eps,
                                                          reinterpret_cast<__half*>(out_hi),
                                                          reinterpret_cast<__half*>(out_lo), out_f32, o_bs, ldo,
                                                          relu);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Small fp32 linear layer for the keypoint encoder's narrow layers (Cin <= 64): y[t, o] = b[o] + sum_c x[t,c] w[o,c].
// The 3->32 and 32->64 layers are <2% of the encoder FLOPs; wider layers go through the tensor-core GEMM.
__global__ void small_linear_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                    const float* __restrict__ bias, float* __restrict__ Y, int ldy, long long rows,
                                    int Cin, int Cout) {
  extern __shared__ float s_w[];  // [Cout][Cin] + [Cout]
  for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) s_w[i] = W[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) s_w[Cout * Cin + i] = bias[i];
  __syncthreads();
  const long long total = rows * Cout;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx / Cout;
    const int o = (int)(idx - t * Cout);
    const float* x = X + t * ldx;
    const float* w = s_w + o * Cin;
    float acc = s_w[Cout * Cin + o];
    for (int cidx = 0; cidx < Cin; ++cidx) acc = fmaf(x[cidx], w[cidx], acc);
    Y[t * ldy + o] = acc;
  }
}

int launch_small_linear(const float* X, int ldx, const float* W, const float* bias, float* Y, int ldy, long long rows,
                        int Cin, int Cout, cudaStream_t st) {
  if (rows == 0) return 0;
  const size_t smem = (size_t)(Cout * Cin + Cout) * sizeof(float);
  IMP_REQUIRE(smem <= 48 * 1024, "small_linear: weight does not fit in shared memory");
  const long long total = rows * Cout;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
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
