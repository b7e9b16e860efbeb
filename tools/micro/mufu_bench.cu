// Microbenchmark: what the softmax inner loop of the attention kernel can sustain on one SM sub-partition.
// Prints cycles per 32-lane exponential ("per warp-element") for several instruction mixes and warps per sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/mufu_bench tools/micro/mufu_bench.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pk(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t ph2(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

// degree-3 polynomial 2^x on the FMA pipe, two values at a time (x <= 0): n = round(x), f = x - n in [-0.5, 0.5]
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& y0, float& y1) {
  const uint64_t magic = pk(12582912.f, 12582912.f);
  x0 = fmaxf(x0, -126.f);
  x1 = fmaxf(x1, -126.f);
  const uint64_t x = pk(x0, x1);
  const uint64_t t = add2(x, magic);                     // integer part in the low mantissa bits
  const uint64_t nf = add2(t, pk(-12582912.f, -12582912.f));
  const uint64_t f = add2(x, mul2(nf, pk(-1.f, -1.f)));  // could be one FFMA2
  uint64_t pcoef = fma2(f, pk(0.0555041f, 0.0555041f), pk(0.2402265f, 0.2402265f));
  pcoef = fma2(pcoef, f, pk(0.6931472f, 0.6931472f));
  pcoef = fma2(pcoef, f, pk(1.f, 1.f));
  float p0, p1, t0, t1;
  upk(pcoef, p0, p1);
  upk(t, t0, t1);
  y0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  y1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int iters, float a, float b) {
  float x[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) x[i] = -0.001f * (threadIdx.x + i);
  float acc = 0.f;
  uint64_t acc2 = pk(0.f, 0.f);
  uint32_t accp = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // MUFU only
#pragma unroll
      for (int i = 0; i < 64; ++i) x[i] = ex2(x[i]) - 1.5f * 0.f + b;  // b = -1: keeps values in range; 1 FADD per MUFU
    } else if (MODE == 1) {  // old kernel mix: FFMA + MUFU + FADD + F2FP(pair)
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const float p0 = ex2(fmaf(x[i], a, b)), p1 = ex2(fmaf(x[i + 1], a, b));
        acc += p0 + p1;
        accp += ph2(p0, p1);
        x[i] = p0;
        x[i + 1] = p1;
      }
    } else if (MODE == 2) {  // packed mix: FFMA2 + 2 MUFU + FADD2 + F2FP per pair
      const uint64_t a2 = pk(a, a), b2 = pk(b, b);
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        float y0, y1;
        upk(fma2(pk(x[i], x[i + 1]), a2, b2), y0, y1);
        const float p0 = ex2(y0), p1 = ex2(y1);
        acc2 = add2(acc2, pk(p0, p1));
        accp += ph2(p0, p1);
        x[i] = p0;
        x[i + 1] = p1;
      }
    } else if (MODE == 3 || MODE == 4 || MODE == 5) {  // packed mix with 1/4 (3), 1/2 (4), 1/8 (5) of the pairs on the FMA pipe
      const uint64_t a2 = pk(a, a), b2 = pk(b, b);
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        float y0, y1, p0, p1;
        upk(fma2(pk(x[i], x[i + 1]), a2, b2), y0, y1);
        const bool poly = MODE == 3 ? ((i & 6) == 6) : (MODE == 4 ? ((i & 2) == 2) : ((i & 14) == 14));
        if (poly) {
          exp2_poly2(y0, y1, p0, p1);
        } else {
          p0 = ex2(y0);
          p1 = ex2(y1);
        }
        acc2 = add2(acc2, pk(p0, p1));
        accp += ph2(p0, p1);
        x[i] = p0;
        x[i + 1] = p1;
      }
    } else if (MODE == 6) {  // MUFU + F2FP only
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const float p0 = ex2(x[i]), p1 = ex2(x[i + 1]);
        accp += ph2(p0, p1);
        x[i] = -p0;
        x[i + 1] = -p1;
      }
    }
  }
  const long long t1 = clock64();
  float s0, s1;
  upk(acc2, s0, s1);
  float r = acc + s0 + s1 + __uint_as_float(accp & 0x3fffffff);
#pragma unroll
  for (int i = 0; i < 64; ++i) r += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, long long* cyc) {
  for (int warps_per_smsp : {1, 2, 4}) {
    const int threads = 128 * warps_per_smsp, iters = 2000, blocks = 148;
    k<MODE><<<blocks, threads>>>(out, cyc, iters, 0.18f, -1.f);
    cudaDeviceSynchronize();
    k<MODE><<<blocks, threads>>>(out, cyc, iters, 0.18f, -1.f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    // per sub-partition: warps_per_smsp warps x iters x 64 exponentials
    printf("%-46s warps/SMSP %d : %.2f cycles per warp-exponential\n", name, warps_per_smsp, (double)mx / ((double)iters * 64 * warps_per_smsp));
  }
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  run<0>("MUFU.EX2 + FADD", out, cyc);
  run<6>("MUFU.EX2 + F2FP/2", out, cyc);
  run<1>("FFMA + MUFU + FADD + F2FP/2 (round-1 mix)", out, cyc);
  run<2>("FFMA2/2 + MUFU + FADD2/2 + F2FP/2 (packed mix)", out, cyc);
  run<5>("packed mix, 1/8 polynomial", out, cyc);
  run<3>("packed mix, 1/4 polynomial", out, cyc);
  run<4>("packed mix, 1/2 polynomial", out, cyc);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
