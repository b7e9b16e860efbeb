// SuperPoint front-end (reference nets/superpoint.py:97-235) for sm_100a: the step BEFORE the matcher.
// Activations are NHWC fp16 hi/lo planes (the same split-precision format the matcher's GEMMs use: x = hi + lo, ~22 bits),
// so that the 3 x 3 convolutions run on the tensor cores with fp32-level accuracy -- keypoint selection is a chain of hard
// decisions (softmax -> NMS equality tests -> threshold -> top-k) and has to track the reference's fp32 network.
//   conv3x3_kernel      implicit GEMM: M = 128 pixels (an 8 x 16 patch), N = C_out, K = 9 taps x C_in.  The A tile of tap
//                       (dy, dx) is ONE 4-D TMA box {64 channels, 16, 8, 1} at (x0 + dx, y0 + dy): TMA's out-of-bounds zero
//                       fill IS the padding, nothing is ever im2col'ed.  tcgen05.mma kind::f16, three products per K step
//                       (hi.hi + lo.hi + hi.lo), fp32 accumulators in TMEM, bias + ReLU + hi/lo split in the epilogue.
//   the rest            small streaming kernels: first layer (C_in = 1), 2 x 2 max pooling, 65-way softmax + depth-to-space,
//                       the three-round 9 x 9 NMS, ordered compaction + top-k, descriptor normalisation and bilinear sampling.
#include <math.h>

#include "common.h"
#include "ptx.cuh"
#include "../../include/imp_b200.h"

namespace imp {

static constexpr int CV_BW = 16, CV_BH = 8;  // 128-pixel tile = 8 image rows x 16 columns
static constexpr int CV_THREADS = 192;
static constexpr int CV_A_BYTES = 128 * 64 * 2;  // one plane of a 128-pixel x 64-channel tile

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

struct ConvParams {
  int B, H, W, Cin, Cout, relu, pool;
  const float* bias;
  __half *out_hi, *out_lo;
};

// MINB CTAs per SM: the kernel is not persistent, so with one CTA per SM its prologue (barriers, TMEM allocation, first TMA
// round trip) and epilogue are exposed -- ncu on the 64 -> 64 layers: tensor pipe 28 %, issue 14 %.  Where shared memory allows
// (C_out = 64: 2 stages of 48 KB) two CTAs share an SM and cover each other's bubbles.
template <int BN, int STG, int MINB>
__global__ void __launch_bounds__(CV_THREADS, MINB)
conv3x3_kernel(const __grid_constant__ CUtensorMap tm_ah, const __grid_constant__ CUtensorMap tm_al,
               const __grid_constant__ CUtensorMap tm_bh, const __grid_constant__ CUtensorMap tm_bl, const ConvParams p) {
  constexpr int B_BYTES = BN * 64 * 2;
  constexpr int STAGE_BYTES = 2 * CV_A_BYTES + 2 * B_BYTES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STG * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STG;
  uint64_t* acc_full = empty + STG;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_full + 1);
  if (smem_u32(smem) & 1023u) __trap();

  const int warp = threadIdx.x >> 5;
  const int x0 = blockIdx.x * CV_BW, y0 = blockIdx.y * CV_BH, b = blockIdx.z;
  const int cblocks = p.Cin / 64;
  const int nkb = 9 * cblocks;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm_ah);
    tma_prefetch_desc(&tm_al);
    tma_prefetch_desc(&tm_bh);
    tma_prefetch_desc(&tm_bl);
    for (int s = 0; s < STG; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STG;
        const int tap = kb / cblocks, cb = kb - tap * cblocks;
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        mbar_wait(&empty[s], ((kb / STG) & 1) ^ 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        // rows / columns outside the image (the 1-pixel zero padding, and the ragged right / bottom edge) are zero-filled
        tma_load_4d(st, &tm_ah, &full[s], cb * 64, x0 + dx, y0 + dy, b);
        tma_load_4d(st + CV_A_BYTES, &tm_al, &full[s], cb * 64, x0 + dx, y0 + dy, b);
        tma_load_3d(st + 2 * CV_A_BYTES, &tm_bh, &full[s], kb * 64, 0, 0);
        tma_load_3d(st + 2 * CV_A_BYTES + B_BYTES, &tm_bl, &full[s], kb * 64, 0, 0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(FMT_F16, 128, BN, 0, 0);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STG;
        mbar_wait(&full[s], (kb / STG) & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + CV_A_BYTES;
        const uint32_t b_hi = a_hi + 2 * CV_A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t dah = make_smem_desc_sw128(a_hi + kk * 32, 16, 1024), dal = make_smem_desc_sw128(a_lo + kk * 32, 16, 1024);
          const uint64_t dbh = make_smem_desc_sw128(b_hi + kk * 32, 16, 1024), dbl = make_smem_desc_sw128(b_lo + kk * 32, 16, 1024);
          // Two accumulators: the tensor core adds each 16-deep partial product into TMEM with truncation, so the error grows
          // with the number of updates of an accumulator and with its magnitude.  hi.hi (the large terms) gets its own one --
          // a third of the updates -- and the two small cross terms are summed among themselves (measured on the full
          // network: relative score error 3.2e-5 with one accumulator).
          umma_f16_ss(tmem_base, dah, dbh, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          umma_f16_ss(tmem_base + BN, dal, dbh, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          umma_f16_ss(tmem_base + BN, dah, dbl, idesc, 1u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane_id();
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int py = y0 + row / CV_BW, px = x0 + row % CV_BW;
    // pool: nn.MaxPool2d(2, 2) fused (nets/superpoint.py:189,192,195).  A warp holds two image rows of 16 pixels, so the 2 x 2
    // partners of a pixel are lanes ^ 1 and ^ 16; the lane of the even / even pixel writes the pooled one (a trailing odd row
    // or column is dropped, as MaxPool2d does).  Same planes as pooling after the split: rounding is monotone.
    const bool ok = p.pool ? ((row & 1) == 0 && (row & 16) == 0 && py + 1 < p.H && px + 1 < p.W) : (py < p.H && px < p.W);
    const long long pix = p.pool ? ((long long)b * (p.H / 2) + py / 2) * (p.W / 2) + px / 2 : ((long long)b * p.H + py) * p.W + px;
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int cb = 0; cb < BN; cb += 32) {
      uint32_t r[32], r2[32];
      tmem_ld_x32(tmem_base + lane_off + cb, r);
      tmem_ld_x32(tmem_base + lane_off + BN + cb, r2);
      tmem_wait_ld();
      float v[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        v[c] = (__uint_as_float(r[c]) + __uint_as_float(r2[c])) + __ldg(p.bias + cb + c);
        if (p.relu) v[c] = fmaxf(v[c], 0.f);
      }
      if (p.pool) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          v[c] = fmaxf(v[c], __shfl_xor_sync(0xffffffffu, v[c], 1));
          v[c] = fmaxf(v[c], __shfl_xor_sync(0xffffffffu, v[c], 16));
        }
      }
      if (ok) {
        __half* oh = p.out_hi + pix * p.Cout + cb;
        __half* ol = p.out_lo + pix * p.Cout + cb;
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            __half h0, l0, h1, l1;
            split_f16x2(v[c + 2 * u], h0, l0);
            split_f16x2(v[c + 2 * u + 1], h1, l1);
            __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
            hi[u] = *reinterpret_cast<uint32_t*>(&hh);
            lo[u] = *reinterpret_cast<uint32_t*>(&ll);
          }
          *reinterpret_cast<uint4*>(oh + c) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(ol + c) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 4-D fp16 tensor map over an NHWC activation plane: {C, W, H, B}, box {64, 16, 8, 1}, 128-byte swizzle, zero OOB fill
static int make_tmap_nhwc(CUtensorMap* out, const void* base, int C, int W, int H, int B) {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  IMP_REQUIRE(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess &&
                  q == cudaDriverEntryPointSuccess,
              "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  IMP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && C % 8 == 0, "NHWC tensor map: 16-byte alignment");
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, CV_BW, CV_BH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(fp)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstr, box,
                                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IMP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (NHWC %d x %d x %d x %d) failed (%d)", B, H, W, C, (int)r);
  return 0;
}

template <int BN, int STG, int MINB>
static int launch_conv3x3_impl(const imp_sp_conv_args& a, cudaStream_t st) {
  CUtensorMap tah, tal, tbh, tbl;
  if (make_tmap_nhwc(&tah, a.in_hi, a.Cin, a.W, a.H, a.B)) return 3;
  if (make_tmap_nhwc(&tal, a.in_lo, a.Cin, a.W, a.H, a.B)) return 3;
  const int K = 9 * a.Cin;
  if (make_tmap_f16_3d(&tbh, a.w_hi, K, a.Cout, 1, K, 0, 64, BN)) return 3;
  if (make_tmap_f16_3d(&tbl, a.w_lo, K, a.Cout, 1, K, 0, 64, BN)) return 3;
  ConvParams p;
  p.B = a.B;
  p.H = a.H;
  p.W = a.W;
  p.Cin = a.Cin;
  p.Cout = a.Cout;
  p.relu = a.relu;
  p.pool = a.pool;
  p.bias = a.bias;
  p.out_hi = reinterpret_cast<__half*>(a.out_hi);
  p.out_lo = reinterpret_cast<__half*>(a.out_lo);
  const size_t smem = STG * (2 * CV_A_BYTES + 2 * BN * 128) + 256;
  auto kern = conv3x3_kernel<BN, STG, MINB>;
  static DeviceOnce configured;
  if (configured.first()) {
    IMP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IMP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  dim3 grid((a.W + CV_BW - 1) / CV_BW, (a.H + CV_BH - 1) / CV_BH, a.B);
  kern<<<grid, CV_THREADS, smem, st>>>(tah, tal, tbh, tbl, p);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_sp_conv3x3(const imp_sp_conv_args& a, cudaStream_t st) {
  IMP_REQUIRE(a.B > 0 && a.H > 0 && a.W > 0, "sp_conv3x3: empty image");
  IMP_REQUIRE(a.Cin % 64 == 0 && a.Cin >= 64, "sp_conv3x3: C_in must be a multiple of 64 (the first layer has its own kernel)");
  switch (a.Cout) {
    case 64: return launch_conv3x3_impl<64, 2, 2>(a, st);
    case 128: return launch_conv3x3_impl<128, 3, 1>(a, st);
    case 256: return launch_conv3x3_impl<256, 2, 1>(a, st);
    default: IMP_REQUIRE(false, "sp_conv3x3: C_out must be 64, 128 or 256 (got %d)", a.Cout);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// conv1a (nets/superpoint.py:125, 1 -> 64 channels): K = 9, far too thin for the tensor cores.  A thread owns one pixel and
// 8 output channels; fp32 FMA in the reference's accumulation order does not matter at 9 terms.
__device__ __forceinline__ uint64_t sp_pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void sp_unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t sp_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__global__ void __launch_bounds__(256) conv1a_kernel(const float* __restrict__ img, const float* __restrict__ w /*[64][9]*/,
                                                     const float* __restrict__ bias, __half* __restrict__ out_hi,
                                                     __half* __restrict__ out_lo, int B, int H, int W) {
  // A thread owns 4 output channels (g*4 .. g*4+3) of one pixel at a time; 16 threads share a pixel.  Their 36 weights live
  // in registers as 18 packed fp32x2 pairs (channels 2u, 2u+1): one FFMA2 does the tap of two channels (same bits as two
  // FFMAs).  4 rather than 8 channels per thread keeps the kernel at ~64 registers: it is latency-bound (9 dependent-free
  // loads, then compute, then two stores per pixel), so resident warps matter more than per-thread reuse.
  const int g = threadIdx.x & 15;
  uint64_t w2[2][9];
  float br[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) br[e] = bias[g * 4 + e];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) w2[u][tp] = sp_pack2(w[(g * 4 + 2 * u) * 9 + tp], w[(g * 4 + 2 * u + 1) * 9 + tp]);
  const unsigned npix = (unsigned)B * H * W;  // < 2^31 (checked by the launcher): 32-bit index arithmetic
  for (unsigned pix = blockIdx.x * 16u + (threadIdx.x >> 4); pix < npix; pix += gridDim.x * 16u) {
    const unsigned rowi = pix / (unsigned)W;
    const int x = (int)(pix - rowi * (unsigned)W), y = (int)(rowi % (unsigned)H);
    const float* c = img + pix;
    float v[9];
    if (x > 0 && x < W - 1 && y > 0 && y < H - 1) {  // interior pixel: no bounds checks
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) v[tp] = __ldg(c + (tp / 3 - 1) * W + (tp % 3 - 1));
    } else {
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        const int yy = y + tp / 3 - 1, xx = x + tp % 3 - 1;
        v[tp] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(c + (tp / 3 - 1) * W + (tp % 3 - 1)) : 0.f;
      }
    }
    uint32_t hi[2], lo[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      uint64_t acc = sp_pack2(0.f, 0.f);
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) acc = sp_fma2(sp_pack2(v[tp], v[tp]), w2[u][tp], acc);
      float o[2];
      sp_unpack2(acc, o[0], o[1]);
      o[0] = fmaxf(o[0] + br[2 * u], 0.f);
      o[1] = fmaxf(o[1] + br[2 * u + 1], 0.f);
      __half h0, l0, h1, l1;
      split_f16x2(o[0], h0, l0);
      split_f16x2(o[1], h1, l1);
      __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
      hi[u] = *reinterpret_cast<uint32_t*>(&hh);
      lo[u] = *reinterpret_cast<uint32_t*>(&ll);
    }
    *reinterpret_cast<uint2*>(out_hi + (size_t)pix * 64 + g * 4) = make_uint2(hi[0], hi[1]);
    *reinterpret_cast<uint2*>(out_lo + (size_t)pix * 64 + g * 4) = make_uint2(lo[0], lo[1]);
  }
}

int launch_sp_conv1a(const float* img, const float* w, const float* bias, void* out_hi, void* out_lo, int B, int H, int W,
                     cudaStream_t st) {
  const long long npix = (long long)B * H * W;
  IMP_REQUIRE(npix > 0 && npix < (1ll << 31) - (1ll << 26), "sp_conv1a: 0 < B * H * W < 2^31");
  const long long want = (npix + 16 * 16 - 1) / (16 * 16);  // ~16 pixels per thread
  const int grid = (int)(want < 1 ? 1 : (want > 65535 * 16 ? 65535 * 16 : want));
  conv1a_kernel<<<grid, 256, 0, st>>>(img, w, bias, reinterpret_cast<__half*>(out_hi), reinterpret_cast<__half*>(out_lo), B, H, W);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

// 2 x 2 / stride 2 max pooling on hi/lo planes (nn.MaxPool2d(2, 2), nets/superpoint.py:122: odd trailing rows / columns are
// dropped).  max over the reconstructed values hi + lo; re-splitting the winner reproduces its planes exactly.
__global__ void maxpool2_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, __half* __restrict__ out_hi,
                                __half* __restrict__ out_lo, int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, G = C / 8;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)B * Ho * Wo * G) return;
  const int g = t % G;
  const long long op = t / G;
  const int xo = op % Wo, yo = (op / Wo) % Ho, b = op / ((long long)Wo * Ho);
  float best[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) best[e] = -INFINITY;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const long long ip = ((long long)b * H + 2 * yo + (q >> 1)) * W + 2 * xo + (q & 1);
    const uint4 h = *reinterpret_cast<const uint4*>(in_hi + ip * C + g * 8);
    const uint4 l = *reinterpret_cast<const uint4*>(in_lo + ip * C + g * 8);
    const __half2* hp = reinterpret_cast<const __half2*>(&h);
    const __half2* lp = reinterpret_cast<const __half2*>(&l);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 hf = __half22float2(hp[u]), lf = __half22float2(lp[u]);
      best[2 * u] = fmaxf(best[2 * u], hf.x + lf.x);
      best[2 * u + 1] = fmaxf(best[2 * u + 1], hf.y + lf.y);
    }
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    __half h0, l0, h1, l1;
    split_f16x2(best[2 * u], h0, l0);
    split_f16x2(best[2 * u + 1], h1, l1);
    __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
    hi[u] = *reinterpret_cast<uint32_t*>(&hh);
    lo[u] = *reinterpret_cast<uint32_t*>(&ll);
  }
  *reinterpret_cast<uint4*>(out_hi + op * C + g * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(out_lo + op * C + g * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

int launch_sp_maxpool2(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int B, int H, int W, int C, cudaStream_t st) {
  IMP_REQUIRE(C % 8 == 0 && H >= 2 && W >= 2, "sp_maxpool2: C %% 8 == 0 and an image of at least 2 x 2");
  const long long n = (long long)B * (H / 2) * (W / 2) * (C / 8);
  maxpool2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const __half*>(in_hi), reinterpret_cast<const __half*>(in_lo),
                                                              reinterpret_cast<__half*>(out_hi), reinterpret_cast<__half*>(out_lo), B, H, W, C);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

// Dense keypoint scores (nets/superpoint.py:199-203): softmax over the 65 logits of a coarse cell, drop the dust-bin channel,
// depth-to-space: channel c of cell (h, w) -> pixel (8 h + c / 8, 8 w + c % 8).  One warp per cell.
__global__ void scores_kernel(const float* __restrict__ logits, int ld, float* __restrict__ scores, int B, int Hc, int Wc) {
  const long long cell = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (cell >= (long long)B * Hc * Wc) return;
  const float* l = logits + cell * ld;
  const float a0 = l[lane], a1 = l[lane + 32], a2 = lane == 0 ? l[64] : -INFINITY;
  const float m = warp_max(fmaxf(fmaxf(a0, a1), a2));
  const float e0 = expf(a0 - m), e1 = expf(a1 - m), e2 = lane == 0 ? expf(a2 - m) : 0.f;
  const float s = warp_sum(e0 + e1 + e2);
  const int w = cell % Wc, h = (cell / Wc) % Hc, b = cell / ((long long)Wc * Hc);
  float* o = scores + ((long long)b * Hc * 8 + h * 8) * (Wc * 8) + w * 8;
  o[(long long)(lane >> 3) * (Wc * 8) + (lane & 7)] = e0 / s;
  o[(long long)((lane >> 3) + 4) * (Wc * 8) + (lane & 7)] = e1 / s;
}

int launch_sp_scores(const float* logits, int ld, float* scores, int B, int Hc, int Wc, cudaStream_t st) {
  const long long n = (long long)B * Hc * Wc * 32;
  scores_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(logits, ld, scores, B, Hc, Wc);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// simple_nms (nets/superpoint.py:50-66): (2 r + 1)^2 max pooling (stride 1, -inf padding) three times over.  One generic
// tiled kernel; `mode` selects what is pooled and what is written:
//   0: mask = (s == pool(s))
//   1: supp = pool(mask) > 0
//   2: t = supp ? 0 : s;  mask |= (t == pool(t)) & !supp
static constexpr int NMS_T = 32;
__global__ void nms_pool_kernel(const float* __restrict__ s, uint8_t* __restrict__ mask, uint8_t* __restrict__ supp, int H, int W,
                                int r, int mode) {
  extern __shared__ float tile[];  // (NMS_T + 2r)^2 values, then (NMS_T + 2r) x NMS_T row maxima
  const int TW = NMS_T + 2 * r;
  float* rowmax = tile + TW * TW;
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * NMS_T, y0 = blockIdx.y * NMS_T;
  const long long base = (long long)b * H * W;
  for (int i = threadIdx.x; i < TW * TW; i += blockDim.x) {
    const int ty = i / TW, tx = i - ty * TW;
    const int y = y0 + ty - r, x = x0 + tx - r;
    float v = -INFINITY;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const long long q = base + (long long)y * W + x;
      v = mode == 0 ? s[q] : (mode == 1 ? (mask[q] ? 1.f : 0.f) : (supp[q] ? 0.f : s[q]));
    }
    tile[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TW * NMS_T; i += blockDim.x) {  // horizontal pass
    const int ty = i / NMS_T, tx = i - ty * NMS_T;
    float m = -INFINITY;
    for (int d = 0; d <= 2 * r; ++d) m = fmaxf(m, tile[ty * TW + tx + d]);
    rowmax[i] = m;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NMS_T * NMS_T; i += blockDim.x) {  // vertical pass + decision
    const int ty = i / NMS_T, tx = i - ty * NMS_T;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;
    float m = -INFINITY;
    for (int d = 0; d <= 2 * r; ++d) m = fmaxf(m, rowmax[(ty + d) * NMS_T + tx]);
    const float c = tile[(ty + r) * TW + tx + r];
    const long long q = base + (long long)y * W + x;
    if (mode == 0) {
      mask[q] = c == m;
    } else if (mode == 1) {
      supp[q] = m > 0.f;
    } else {
      if (c == m && !supp[q]) mask[q] = 1;
    }
  }
}

int launch_sp_nms(const float* scores, uint8_t* mask, uint8_t* supp, int B, int H, int W, int radius, cudaStream_t st) {
  IMP_REQUIRE(radius >= 0 && radius <= 16, "sp_nms: radius 0..16");
  const int TW = NMS_T + 2 * radius;
  const size_t smem = (size_t)(TW * TW + TW * NMS_T) * 4;
  dim3 grid((W + NMS_T - 1) / NMS_T, (H + NMS_T - 1) / NMS_T, B);
  nms_pool_kernel<<<grid, 256, smem, st>>>(scores, mask, supp, H, W, radius, 0);
  for (int it = 0; it < 2; ++it) {
    nms_pool_kernel<<<grid, 256, smem, st>>>(scores, mask, supp, H, W, radius, 1);
    nms_pool_kernel<<<grid, 256, smem, st>>>(scores, mask, supp, H, W, radius, 2);
  }
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Keypoint extraction (nets/superpoint.py:206-222): nonzero(nms > threshold) in row-major order, border removal, top-k.
__device__ __forceinline__ bool sp_is_kpt(const float* s, const uint8_t* mask, long long q, int y, int x, int H, int W, float thr,
                                          int border) {
  return mask[q] && s[q] > thr && y >= border && y < H - border && x >= border && x < W - border;
}
__global__ void sel_count_kernel(const float* __restrict__ s, const uint8_t* __restrict__ mask, int* __restrict__ rowcnt, int H, int W,
                                 float thr, int border) {
  const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (y >= H) return;
  int n = 0;
  for (int x = lane; x < W; x += 32) n += sp_is_kpt(s, mask, (long long)y * W + x, y, x, H, W, thr, border);
  n = (int)warp_sum((float)n);
  if (lane == 0) rowcnt[y] = n;
}
__global__ void sel_scan_kernel(const int* __restrict__ rowcnt, int* __restrict__ rowoff, int* __restrict__ total, int H) {
  __shared__ int part[1024];
  const int per = (H + 1023) / 1024;
  int acc = 0;
  for (int i = 0; i < per; ++i) {
    const int y = threadIdx.x * per + i;
    if (y < H) acc += rowcnt[y];
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = part[threadIdx.x] - acc;
  for (int i = 0; i < per; ++i) {
    const int y = threadIdx.x * per + i;
    if (y < H) {
      rowoff[y] = run;
      run += rowcnt[y];
    }
  }
  if (threadIdx.x == 1023) *total = part[1023];
}
__global__ void sel_write_kernel(const float* __restrict__ s, const uint8_t* __restrict__ mask, const int* __restrict__ rowoff,
                                 int* __restrict__ cand_yx, float* __restrict__ cand_score, int H, int W, float thr, int border, int cap) {
  const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (y >= H) return;
  int off = rowoff[y];
  for (int x0 = 0; x0 < W; x0 += 32) {
    const int x = x0 + lane;
    const bool k = x < W && sp_is_kpt(s, mask, (long long)y * W + x, y, x, H, W, thr, border);
    const unsigned bal = __ballot_sync(0xffffffffu, k);
    if (k) {
      const int o = off + __popc(bal & ((1u << lane) - 1));
      if (o < cap) {
        cand_yx[2 * o] = y;
        cand_yx[2 * o + 1] = x;
        cand_score[o] = s[(long long)y * W + x];
      }
    }
    off += __popc(bal);
  }
}
// top-k (torch.topk: sorted by descending score; equal scores: lower row-major index first) or, when there are at most k
// candidates, all of them in row-major order (top_k_keypoints returns its input unchanged, nets/superpoint.py:79-80).
// One CTA.  Keys are 64 bit (monotone score bits | inverted index), hence distinct.  k <= 4096 (the usual 1000-4000 keypoints):
// 8-pass radix select of the k-th largest key over the candidates (tens of thousands on a 1600-pixel image), gather of the
// k keys >= it into shared memory, bitonic sort there.  Larger k: bitonic sort of all keys in a global scratch buffer.
static constexpr int SEL_SMEM_K = 4096;
__device__ __forceinline__ unsigned long long sel_key(const float* cand_score, int i) {
  return ((unsigned long long)__float_as_uint(cand_score[i]) << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)i);  // scores > 0
}
template <typename KeyArray>
__device__ __forceinline__ void bitonic_desc(KeyArray keys, int np2) {
  for (int size = 2; size <= np2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < np2 / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));  // index of the lower element of pair i
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;  // descending blocks first -> the whole array ends up descending
        const unsigned long long a = keys[lo], b2 = keys[hi];
        if ((a < b2) == desc) {
          keys[lo] = b2;
          keys[hi] = a;
        }
      }
      __syncthreads();
    }
  }
}
__global__ void __launch_bounds__(1024) sel_topk_kernel(const int* __restrict__ cand_yx, const float* __restrict__ cand_score,
                                                         const int* __restrict__ total, unsigned long long* __restrict__ keys, int cap,
                                                         int k, float* __restrict__ kpts_xy, float* __restrict__ kscores,
                                                         int* __restrict__ n_out) {
  __shared__ unsigned long long sel[SEL_SMEM_K];
  __shared__ int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_remaining, s_slot;
  const int n = min(*total, cap);
  const bool all = k < 0 || n <= k;
  const int m = all ? n : k;
  const unsigned long long* sorted = keys;
  if (!all && k <= SEL_SMEM_K) {
    if (threadIdx.x == 0) {
      s_prefix = 0ull;
      s_remaining = k;
      s_slot = 0;
    }
    unsigned long long mask = 0ull;
    for (int shift = 56; shift >= 0; shift -= 8) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long key = sel_key(cand_score, i);
        if ((key & mask) == prefix) atomicAdd(&hist[(int)((key >> shift) & 255ull)], 1);
      }
      __syncthreads();
      if (threadIdx.x == 0) {  // the digit of the k-th largest key among those that share the prefix
        int rem = s_remaining, b = 255;
        while (b > 0 && hist[b] < rem) rem -= hist[b--];
        s_remaining = rem;
        s_prefix = prefix | ((unsigned long long)b << shift);
      }
      mask |= 255ull << shift;
      __syncthreads();
    }
    const unsigned long long kth = s_prefix;  // exactly k keys are >= kth
    int np2 = 1;
    while (np2 < k) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += blockDim.x) sel[i] = 0ull;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned long long key = sel_key(cand_score, i);
      if (key >= kth) sel[atomicAdd(&s_slot, 1)] = key;
    }
    __syncthreads();
    bitonic_desc(sel, np2);
    sorted = sel;
  } else if (!all) {
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += blockDim.x) keys[i] = i < n ? sel_key(cand_score, i) : 0ull;
    __syncthreads();
    bitonic_desc(keys, np2);
  }
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const int src = all ? i : (int)(0xFFFFFFFFu - (unsigned)(sorted[i] & 0xFFFFFFFFull));
    kpts_xy[2 * i] = (float)cand_yx[2 * src + 1];  // (h, w) -> (x, y), nets/superpoint.py:225
    kpts_xy[2 * i + 1] = (float)cand_yx[2 * src];
    kscores[i] = cand_score[src];
  }
  // fixed-capacity consumers (SuperPoint.detect_padded -> the matcher's 'n_keypoints' convention) want zero padding up to k
  for (int i = m + threadIdx.x; i < min(k, cap); i += blockDim.x) {
    kpts_xy[2 * i] = 0.f;
    kpts_xy[2 * i + 1] = 0.f;
    kscores[i] = 0.f;
  }
  if (threadIdx.x == 0) *n_out = m;
}

int launch_sp_select(const imp_sp_select_args& a, cudaStream_t st) {
  IMP_REQUIRE(a.H > 0 && a.W > 0 && a.cap > 0, "sp_select: empty problem");
  IMP_REQUIRE(a.H <= 1024 * 64, "sp_select: image too tall");
  const int wpb = 8;
  sel_count_kernel<<<(a.H + wpb - 1) / wpb, wpb * 32, 0, st>>>(a.scores, a.mask, a.rowcnt, a.H, a.W, a.threshold, a.border);
  sel_scan_kernel<<<1, 1024, 0, st>>>(a.rowcnt, a.rowoff, a.total, a.H);
  sel_write_kernel<<<(a.H + wpb - 1) / wpb, wpb * 32, 0, st>>>(a.scores, a.mask, a.rowoff, a.cand_yx, a.cand_score, a.H, a.W,
                                                             a.threshold, a.border, a.cap);
  sel_topk_kernel<<<1, 1024, 0, st>>>(a.cand_yx, a.cand_score, a.total, reinterpret_cast<unsigned long long*>(a.keys), a.cap,
                                      a.max_keypoints, a.kpts_xy, a.kscores, a.n_out);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Descriptors: L2 normalisation of the dense map (nets/superpoint.py:229), bilinear sampling at the keypoints with the
// reference's coordinate convention + a second normalisation (sample_descriptors, :83-95).  One warp per row / keypoint.
__global__ void l2norm_rows_kernel(float* __restrict__ x, long long rows, int ld) {
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float4* p = reinterpret_cast<float4*>(x + r * ld);
  float4 a = p[lane], b = p[lane + 32];
  float ss = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  a.x *= inv, a.y *= inv, a.z *= inv, a.w *= inv, b.x *= inv, b.y *= inv, b.z *= inv, b.w *= inv;
  p[lane] = a;
  p[lane + 32] = b;
}
int launch_sp_l2norm_rows(float* x, long long rows, int ld, cudaStream_t st) {
  IMP_REQUIRE(ld >= 256 && ld % 4 == 0, "sp_l2norm_rows: 256-channel rows");
  l2norm_rows_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(x, rows, ld);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void sample_desc_kernel(const float* __restrict__ dmap /*[Hc*Wc][256]*/, const float* __restrict__ kpts_xy,
                                   const int* __restrict__ n_kpts, float* __restrict__ out /*[K][256]*/, int Hc, int Wc, int max_k) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n = n_kpts ? min(*n_kpts, max_k) : max_k;
  if (k >= max_k) return;
  if (k >= n) {  // zero padding behind the real keypoints (fixed-capacity output)
    float4* o = reinterpret_cast<float4*>(out + (long long)k * 256);
    o[lane] = o[lane + 32] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float s = 8.f;
  // keypoints - s/2 + 0.5, / (w s - s/2 - 0.5), * 2 - 1, then grid_sample
  float gx = (kpts_xy[2 * k] - s / 2 + 0.5f) / (Wc * s - s / 2 - 0.5f);
  float gy = (kpts_xy[2 * k + 1] - s / 2 + 0.5f) / (Hc * s - s / 2 - 0.5f);
  gx = gx * 2.f - 1.f;
  gy = gy * 2.f - 1.f;
  // align_corners = False (see include/imp_b200.h): ((g + 1) * size - 1) / 2
  const float ix = ((gx + 1.f) * Wc - 1.f) / 2.f, iy = ((gy + 1.f) * Hc - 1.f) / 2.f;
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float wx1 = ix - fx, wx0 = 1.f - wx1, wy1 = iy - fy, wy0 = 1.f - wy1;
  float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int xx = x0 + (q & 1), yy = y0 + (q >> 1);
    const float w = ((q & 1) ? wx1 : wx0) * ((q >> 1) ? wy1 : wy0);
    if (xx < 0 || xx >= Wc || yy < 0 || yy >= Hc) continue;  // zero padding
    const float4* p = reinterpret_cast<const float4*>(dmap + ((long long)yy * Wc + xx) * 256);
    const float4 a = p[lane], b = p[lane + 32];
    acc0.x = fmaf(a.x, w, acc0.x), acc0.y = fmaf(a.y, w, acc0.y), acc0.z = fmaf(a.z, w, acc0.z), acc0.w = fmaf(a.w, w, acc0.w);
    acc1.x = fmaf(b.x, w, acc1.x), acc1.y = fmaf(b.y, w, acc1.y), acc1.z = fmaf(b.z, w, acc1.z), acc1.w = fmaf(b.w, w, acc1.w);
  }
  float ss = acc0.x * acc0.x + acc0.y * acc0.y + acc0.z * acc0.z + acc0.w * acc0.w + acc1.x * acc1.x + acc1.y * acc1.y +
             acc1.z * acc1.z + acc1.w * acc1.w;
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  acc0.x *= inv, acc0.y *= inv, acc0.z *= inv, acc0.w *= inv, acc1.x *= inv, acc1.y *= inv, acc1.z *= inv, acc1.w *= inv;
  float4* o = reinterpret_cast<float4*>(out + (long long)k * 256);
  o[lane] = acc0;
  o[lane + 32] = acc1;
}
int launch_sp_sample_descriptors(const float* dmap, const float* kpts_xy, const int* n_kpts, float* out, int Hc, int Wc, int max_k,
                                 cudaStream_t st) {
  if (max_k <= 0) return 0;
  sample_desc_kernel<<<(max_k * 32 + 255) / 256, 256, 0, st>>>(dmap, kpts_xy, n_kpts, out, Hc, Wc, max_k);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace imp
