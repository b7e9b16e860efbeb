// tcgen05 split-precision GEMM (see gemm.cuh).  Persistent: one CTA per SM walks over 128 x BN output tiles.
//   warp 0  : TMA producer  (hi/lo planes of A and B, 64-wide K slabs, 128B swizzle; the smem ring runs across tiles)
//   warp 1  : UMMA issuer   (one elected lane; two accumulators of 128 lanes x BN TMEM columns, ping-pong)
//   warp 2  : TMEM allocator
//   warps 4-7: epilogue     (tcgen05.ld, bias / residual / hi-lo split, vectorised global stores) -- drains
//             accumulator t while the tensor core already works on tile t+1
#include "gemm.cuh"

#include "common.h"
#include "ptx.cuh"

#include <stdlib.h>

namespace imp {

static constexpr int GEMM_BM = 128;
static constexpr int GEMM_THREADS = 256;

struct GemmKernelParams {
  int M, N, KB1, KB, b_batched, nsplit, out_mode;
  int tiles_n, tiles_m, tiles_total;
  float alpha;
  const float* bias;
  void *out0, *out1;
  const __half *res_hi, *res_lo;
  long long out_row_stride, out_batch_stride;
  // instance-norm statistics of the fp32 output (per image of stat_np rows: column sums / sums of squares over the
  // first stat_ns[img] rows), see GemmArgs
  float2* stat_partial;   // [tiles_m][N]
  float2* stat_straddle;  // [images][4][N]
  const int* stat_ns;
  int stat_np;
  // NORM_A: A = relu((H - mean) * rstd) formed in the kernel from an fp32 H tile (tensor map tm_a_hi) and a_stats
  const float2* a_stats;  // [images][K] (mean, rstd)
  int a_np;               // rows per image
};

// Epilogue of one 128 x BN accumulator tile by the four epilogue warps of a CTA (warp quarter q, TMEM lane = row):
// bias / alpha / residual, output modes, coalesced stores through a warp-private staging tile, optional instance-norm
// statistics (per-slab sums into s_stat / stat_straddle).  `t_acc` = TMEM address of the accumulator's column 0 (lane
// field 0); `wait_full` is invoked once the bias slice is staged and returns when the accumulator is complete.
template <int BN, typename WaitFull>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmKernelParams& p, uint32_t t_acc, float* bias_t, uint8_t* s_stage,
                                                   float2* s_stat, int m0, int n0, int z, WaitFull wait_full) {
  const int q = (threadIdx.x >> 5) & 3;  // TMEM lane quarter this warp may touch
  const int row = q * 32 + lane_id();
  const long long gm = m0 + row;
  const bool row_ok = gm < p.M;
  // Instance-norm statistics (stats mode): rows of this warp's 32-row slab that are valid tokens of the tile's first
  // image (segment A: [sa0, sa1)) and of the image the tile straddles into (segment B: [sb0, sb1)).  Warp-uniform.
  const bool stats = p.stat_partial != nullptr;
  int sa0 = 0, sa1 = 0, sb0 = 0, sb1 = 0, img_b = -1;
  if (stats) {
    const int r0 = m0 + q * 32;                 // first global row of the slab
    const int img_a = m0 / p.stat_np;
    const int end_a = (img_a + 1) * p.stat_np;  // first row of the next image
    sa1 = min(32, img_a * p.stat_np + p.stat_ns[img_a] - r0);
    sa1 = max(sa1, 0);
    if (end_a < m0 + GEMM_BM && end_a < p.M) {  // the tile reaches into image img_a + 1
      img_b = img_a + 1;
      sb0 = max(0, end_a - r0);
      sb1 = min(32, end_a + p.stat_ns[img_b] - r0);
      if (sb1 < sb0) sb1 = sb0;
      if (sb0 >= 32) sb0 = sb1 = 0;
    }
  }
  // stage this tile's bias slice in shared memory (one coalesced load instead of 256 broadcast LDGs per thread)
  for (int j = threadIdx.x - 128; j < BN; j += 128)
    bias_t[j] = (p.bias != nullptr && n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
  asm volatile("bar.sync 2, 128;" ::: "memory");
  wait_full();
  tc_fence_after();
  const uint32_t t_row = t_acc + (static_cast<uint32_t>(q * 32) << 16);
  const long long obase = (long long)z * p.out_batch_stride + gm * p.out_row_stride;
  const int nchunks = min(BN / 32, (p.N - n0 + 31) / 32);
  uint32_t r[32];
  tmem_ld_x32(t_row, r);
  // residual planes of the warp's 32 x 32 sub-tile, prefetched one chunk ahead (coalesced: 8 rows x 64 B per pass)
  const bool resid = p.out_mode == GEMM_OUT_SPLIT_RESID;
  const int rows_ok_w = min(32, p.M - (m0 + q * 32));
  const long long wbase = (long long)z * p.out_batch_stride + (long long)(m0 + q * 32) * p.out_row_stride;
  uint4 rh[4], rl[4];
  auto load_resid = [&](int nc_) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int rr = it * 8 + (lane_id() >> 2), seg = lane_id() & 3;
      rh[it] = make_uint4(0, 0, 0, 0);
      rl[it] = rh[it];
      if (rr < rows_ok_w && nc_ + 32 <= p.N) {
        rh[it] = *reinterpret_cast<const uint4*>(p.res_hi + wbase + nc_ + (long long)rr * p.out_row_stride + seg * 8);
        rl[it] = *reinterpret_cast<const uint4*>(p.res_lo + wbase + nc_ + (long long)rr * p.out_row_stride + seg * 8);
      }
    }
  };
  if (resid) load_resid(n0);
#pragma unroll 1
  for (int c = 0; c < nchunks; ++c) {
    const int nc = n0 + c * 32;
    tmem_wait_ld();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), p.alpha, bias_t[c * 32 + j]);
    if (c + 1 < nchunks) tmem_ld_x32(t_row + (c + 1) * 32, r);  // next chunk's TMEM read overlaps this chunk's stores
    const bool full = nc + 32 <= p.N;
    if (full) {
      // Coalesced path: the warp's 32 x 32 sub-tile goes through a warp-private shared-memory tile so that
      // global memory sees whole 64/128-byte row segments instead of 32 scattered 16-byte pieces per instruction.
      uint8_t* stg = s_stage + q * (32 * 144);
      const int l = lane_id();
      const long long wrow0 = (long long)z * p.out_batch_stride + (long long)(m0 + q * 32) * p.out_row_stride + nc;
      const int rows_ok = min(32, p.M - (m0 + q * 32));  // rows of this warp that exist (may be <= 0)
      if (resid) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {  // the prefetched 8 rows x (64 B hi + 64 B lo) per pass
          const int rr = it * 8 + (l >> 2), seg = l & 3;
          *reinterpret_cast<uint4*>(stg + rr * 144 + seg * 16) = rh[it];
          *reinterpret_cast<uint4*>(stg + rr * 144 + 64 + seg * 16) = rl[it];
        }
        __syncwarp();
        if (c + 1 < nchunks) load_resid(nc + 32);  // next chunk's residual is in flight during this chunk's stores
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 ra = *reinterpret_cast<const uint4*>(stg + l * 144 + j * 16);
          const uint4 rb = *reinterpret_cast<const uint4*>(stg + l * 144 + 64 + j * 16);
          const __half2* ah = reinterpret_cast<const __half2*>(&ra);
          const __half2* bh = reinterpret_cast<const __half2*>(&rb);
#pragma unroll
          for (int t2 = 0; t2 < 4; ++t2) {
            const float2 fa = __half22float2(ah[t2]);
            const float2 fb = __half22float2(bh[t2]);
            v[8 * j + 2 * t2] += fa.x + fb.x;
            v[8 * j + 2 * t2 + 1] += fa.y + fb.y;
          }
        }
        __syncwarp();
      }
      if (p.out_mode == GEMM_OUT_F32) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stg + l * 144 + j * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        if (stats) {
          // lane l owns column nc + l of the 32 x 32 sub-tile sitting in the staging tile (pitch 36 words: conflict-free)
          float s1 = 0.f, s2 = 0.f;
          for (int rr = sa0; rr < sa1; ++rr) {
            const float x = *reinterpret_cast<const float*>(stg + rr * 144 + l * 4);
            s1 += x;
            s2 = fmaf(x, x, s2);
          }
          s_stat[q * BN + c * 32 + l] = make_float2(s1, s2);
          if (img_b >= 0) {
            float t1 = 0.f, t2 = 0.f;
            for (int rr = sb0; rr < sb1; ++rr) {
              const float x = *reinterpret_cast<const float*>(stg + rr * 144 + l * 4);
              t1 += x;
              t2 = fmaf(x, x, t2);
            }
            p.stat_straddle[((long long)img_b * 4 + q) * p.N + nc + l] = make_float2(t1, t2);
          }
        }
        float* o = reinterpret_cast<float*>(p.out0) + wrow0;
#pragma unroll
        for (int it = 0; it < 8; ++it) {  // 4 rows x 128 B per pass
          const int r = it * 4 + (l >> 3), seg = l & 7;
          if (r < rows_ok)
            *reinterpret_cast<float4*>(o + (long long)r * p.out_row_stride + seg * 4) =
                *reinterpret_cast<const float4*>(stg + r * 144 + seg * 16);
        }
      } else if (p.out_mode == GEMM_OUT_F16) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 pk;
          pk.x = pack_half2(v[j], v[j + 1]);
          pk.y = pack_half2(v[j + 2], v[j + 3]);
          pk.z = pack_half2(v[j + 4], v[j + 5]);
          pk.w = pack_half2(v[j + 6], v[j + 7]);
          *reinterpret_cast<uint4*>(stg + l * 144 + j * 2) = pk;
        }
        __syncwarp();
        __half* o = reinterpret_cast<__half*>(p.out0) + wrow0;
#pragma unroll
        for (int it = 0; it < 4; ++it) {  // 8 rows x 64 B per pass
          const int r = it * 8 + (l >> 2), seg = l & 3;
          if (r < rows_ok)
            *reinterpret_cast<uint4*>(o + (long long)r * p.out_row_stride + seg * 8) =
                *reinterpret_cast<const uint4*>(stg + r * 144 + seg * 16);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int t2 = 0; t2 < 4; ++t2) {
            __half h0, l0, h1, l1;
            split_f16x2(v[j + 2 * t2], h0, l0);
            split_f16x2(v[j + 2 * t2 + 1], h1, l1);
            __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
            hi[t2] = *reinterpret_cast<uint32_t*>(&hh);
            lo[t2] = *reinterpret_cast<uint32_t*>(&ll);
          }
          *reinterpret_cast<uint4*>(stg + l * 144 + j * 2) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(stg + l * 144 + 64 + j * 2) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        __syncwarp();
        __half* oh = reinterpret_cast<__half*>(p.out0) + wrow0;
        __half* ol = reinterpret_cast<__half*>(p.out1) + wrow0;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int r = it * 8 + (l >> 2), seg = l & 3;
          if (r < rows_ok) {
            *reinterpret_cast<uint4*>(oh + (long long)r * p.out_row_stride + seg * 8) =
                *reinterpret_cast<const uint4*>(stg + r * 144 + seg * 16);
            *reinterpret_cast<uint4*>(ol + (long long)r * p.out_row_stride + seg * 8) =
                *reinterpret_cast<const uint4*>(stg + r * 144 + 64 + seg * 16);
          }
        }
      }
      __syncwarp();  // staging tile is reused by the next chunk
      continue;
    }
    // ragged last chunk (N not a multiple of 32): per-thread scalar path
    if (!row_ok) continue;
    if (p.out_mode == GEMM_OUT_F32) {
      float* o = reinterpret_cast<float*>(p.out0) + obase + nc;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (nc + j < p.N) o[j] = v[j];
    } else if (p.out_mode == GEMM_OUT_F16) {
      __half* o = reinterpret_cast<__half*>(p.out0) + obase + nc;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (nc + j < p.N) o[j] = __float2half_rn(v[j]);
    } else {
      __half* oh = reinterpret_cast<__half*>(p.out0) + obase + nc;
      __half* ol = reinterpret_cast<__half*>(p.out1) + obase + nc;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (nc + j < p.N) {
          float x = v[j];
          if (p.out_mode == GEMM_OUT_SPLIT_RESID) x += __half2float(p.res_hi[obase + nc + j]) + __half2float(p.res_lo[obase + nc + j]);
          __half h, lw;
          split_f16x2(x, h, lw);
          oh[j] = h;
          ol[j] = lw;
        }
      }
    }
  }
}

// the four slabs of a tile, added in fixed order (deterministic): one (sum, sum of squares) per tile and column
template <int BN>
__device__ __forceinline__ void gemm_epilogue_stats_flush(const GemmKernelParams& p, const float2* s_stat, int m0, int n0) {
  asm volatile("bar.sync 2, 128;" ::: "memory");
  for (int j = threadIdx.x - 128; j < BN && n0 + j < p.N; j += 128) {
    const float2 a0 = s_stat[j], a1 = s_stat[BN + j], a2 = s_stat[2 * BN + j], a3 = s_stat[3 * BN + j];
    p.stat_partial[(long long)(m0 / GEMM_BM) * p.N + n0 + j] =
        make_float2(((a0.x + a1.x) + a2.x) + a3.x, ((a0.y + a1.y) + a2.y) + a3.y);
  }
  // (the next tile writes s_stat only after its bias-staging barrier, which every thread reaches after this loop)
}

// NORM_A transform of one staged K block (64 wide) by 128 threads: row r of the fp32 tile at `st` (two 128-byte-swizzled
// boxes of 32 floats x 128 rows, 16 KB each) becomes relu((x - mean) * rstd) as fp16 hi / lo planes in place.
__device__ __forceinline__ void gemm_norm_a_row(const GemmKernelParams& p, uint8_t* st, const float2* s_ms, int r, int m0,
                                                int img_a, int K, int kb) {
  constexpr int A_BYTES = GEMM_BM * 64 * 2;
  {
    const int which = ((m0 + r) / p.a_np > img_a) ? 1 : 0;
    const float2* stp = s_ms + which * K + kb * 64;
    uint8_t* row_lo = st + r * 128;            // fp32 floats 0..31 of the K block, later the hi plane
    uint8_t* row_hi = st + A_BYTES + r * 128;  // fp32 floats 32..63, later the lo plane
    const int sw = r & 7;
    float x[64];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 a = *reinterpret_cast<const float4*>(row_lo + ((c ^ sw) << 4));
      const float4 b = *reinterpret_cast<const float4*>(row_hi + ((c ^ sw) << 4));
      x[4 * c] = a.x; x[4 * c + 1] = a.y; x[4 * c + 2] = a.z; x[4 * c + 3] = a.w;
      x[32 + 4 * c] = b.x; x[32 + 4 * c + 1] = b.y; x[32 + 4 * c + 2] = b.z; x[32 + 4 * c + 3] = b.w;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {  // output chunk c = K elements 8c .. 8c+7
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 ms = *reinterpret_cast<const float4*>(stp + 8 * c + 2 * e);  // (mean, rstd) x 2, broadcast
        const float y0 = fmaxf((x[8 * c + 2 * e] - ms.x) * ms.y, 0.f);
        const float y1 = fmaxf((x[8 * c + 2 * e + 1] - ms.z) * ms.w, 0.f);
        __half h0, l0, h1, l1;
        split_f16x2(y0, h0, l0);
        split_f16x2(y1, h1, l1);
        __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
        hi[e] = *reinterpret_cast<uint32_t*>(&hh);
        lo[e] = *reinterpret_cast<uint32_t*>(&ll);
      }
      *reinterpret_cast<uint4*>(row_lo + ((c ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(row_hi + ((c ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}
// (mean, rstd) of the (at most two) images a 128-row tile touches -> shared memory [2][K] float2, by 128 threads
__device__ __forceinline__ void gemm_norm_a_stage_stats(const GemmKernelParams& p, float2* s_ms, int tid, int img_a, int K) {
  const int n_img = p.M / p.a_np;
  asm volatile("bar.sync 3, 128;" ::: "memory");  // every thread is done with the previous tile's statistics
  for (int i = tid; i < 2 * K / 2; i += 128) {    // float4 = two (mean, rstd) pairs
    const int which = i / (K / 2), j = i - which * (K / 2);
    const int img = min(img_a + which, n_img - 1);
    reinterpret_cast<float4*>(s_ms)[i] = __ldg(reinterpret_cast<const float4*>(p.a_stats + (long long)img * K) + j);
  }
  asm volatile("bar.sync 3, 128;" ::: "memory");
}

// BK = K elements per pipeline stage = one swizzle span (64 fp16 = 128 B, or 32 fp16 = 64 B).  The mainloop is bound by
// the latency of the TMA fetches, not by their bandwidth, so four 48 KB stages (BK = 32) beat two 96 KB stages (BK = 64).
template <int BN, int BK>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int SBO = 8 * BK * 2;  // bytes between 8-row groups
};

// NORM_A (the second half of the instance norm disappears into the next GEMM): the A operand is not read
// as fp16 planes but as the fp32 pre-norm tile H (two 128-byte-swizzled TMA boxes of 32 floats x 128 rows per 64-wide K
// block, landing exactly where A_hi | A_lo live); warps 2-3 turn every row IN PLACE into relu((H - mean) * rstd) split into
// the hi / lo planes (row r of both fp32 boxes occupies the same bytes as row r of A_hi and A_lo, and the 16-byte chunks
// of both layouts are XOR-swizzled by (r & 7)), fence to the async proxy and signal the issuer.
template <int BN, int STAGES, int BK, bool NORM_A = false>
__global__ void __launch_bounds__(NORM_A ? GEMM_THREADS + 64 : GEMM_THREADS, 1)
gemm_f16split_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                     const __grid_constant__ CUtensorMap tm_a2_hi, const __grid_constant__ CUtensorMap tm_a2_lo,
                     const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                     const GemmKernelParams p) {
  using S = GemmSmem<BN, BK>;
  constexpr int ACC = (2 * BN <= 512) ? 2 : 1;  // accumulator buffers in TMEM
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [ACC]
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC;  // [ACC]
  uint64_t* a_ready_bar = tmem_empty_bar + ACC;    // [STAGES] NORM_A: the A planes of the stage have been produced
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(a_ready_bar + STAGES);
  // [ACC][BN] bias slice of the tile being drained; 16-byte aligned (the staging tiles behind it take uint4 accesses)
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr_smem + 4) + 15) & ~uintptr_t(15));
  // per-epilogue-warp staging tile: 32 rows x 128 B payload, 144 B pitch (conflict-free for 16-byte accesses)
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_bias + ACC * BN);
  float2* s_stat = reinterpret_cast<float2*>(s_stage + 4 * 32 * 144);  // [4 warps][BN] (sum, sum of squares), stats mode only

  const int warp = threadIdx.x >> 5;
  const bool split = p.nsplit == 3;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_b_hi);
    if (split) {
      tma_prefetch_desc(&tm_a_lo);
      tma_prefetch_desc(&tm_b_lo);
    }
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&a_ready_bar[s], 128);
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, ACC * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // tile t -> (n tile fastest, then m tile, then batch): CTAs running side by side share the A tile in L2
  auto tile_coords = [&](int t, int& m0, int& n0, int& z) {
    const int per_z = p.tiles_n * p.tiles_m;
    z = t / per_z;
    const int r = t - z * per_z;
    m0 = (r / p.tiles_n) * GEMM_BM;
    n0 = (r % p.tiles_n) * BN;
  };

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t tx = split ? S::STAGE_BYTES : (S::A_BYTES + S::B_BYTES);
      int it = 0;  // running k-block counter across tiles
      for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x) {
        int m0, n0, z;
        tile_coords(t, m0, n0, z);
        for (int kb = 0; kb < p.KB; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* st = smem + s * S::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], tx);
          const bool seg2 = kb >= p.KB1;
          const int ka = (seg2 ? kb - p.KB1 : kb) * BK;
          if (NORM_A) {  // fp32 H tile: floats ka .. ka+31 and ka+32 .. ka+63 of 128 rows
            tma_load_3d(st, &tm_a_hi, &full_bar[s], ka, m0, z);
            tma_load_3d(st + S::A_BYTES, &tm_a_hi, &full_bar[s], ka + 32, m0, z);
          } else {
            tma_load_3d(st, seg2 ? &tm_a2_hi : &tm_a_hi, &full_bar[s], ka, m0, z);
          }
          tma_load_3d(st + 2 * S::A_BYTES, &tm_b_hi, &full_bar[s], kb * BK, n0, p.b_batched ? z : 0);
          if (split) {
            if (!NORM_A) tma_load_3d(st + S::A_BYTES, seg2 ? &tm_a2_lo : &tm_a_lo, &full_bar[s], ka, m0, z);
            tma_load_3d(st + 2 * S::A_BYTES + S::B_BYTES, &tm_b_lo, &full_bar[s], kb * BK, n0,
                        p.b_batched ? z : 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(FMT_F16, GEMM_BM, BN, 0, 0);
      int it = 0, lt = 0;
      for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x, ++lt) {
        const int a = lt % ACC;
        mbar_wait(&tmem_empty_bar[a], ((lt / ACC) & 1) ^ 1);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        for (int kb = 0; kb < p.KB; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(NORM_A ? &a_ready_bar[s] : &full_bar[s], (it / STAGES) & 1);  // a_ready implies full (B landed too)
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + s * S::STAGE_BYTES);
          const uint32_t a_lo = a_hi + S::A_BYTES;
          const uint32_t b_hi = a_hi + 2 * S::A_BYTES;
          const uint32_t b_lo = b_hi + S::B_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint32_t off = k * 32;  // 16 fp16 along K inside the swizzle span
            constexpr uint32_t LT = (BK == 64) ? 2u : 4u;  // SWIZZLE_128B : SWIZZLE_64B
            const uint64_t dah = make_smem_desc(a_hi + off, 16, S::SBO, LT);
            const uint64_t dbh = make_smem_desc(b_hi + off, 16, S::SBO, LT);
            umma_f16_ss(d_tmem, dah, dbh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (split) {
              const uint64_t dal = make_smem_desc(a_lo + off, 16, S::SBO, LT);
              const uint64_t dbl = make_smem_desc(b_lo + off, 16, S::SBO, LT);
              umma_f16_ss(d_tmem, dal, dbh, idesc, 1u);
              umma_f16_ss(d_tmem, dah, dbl, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs retire
        }
        umma_commit(&tmem_full_bar[a]);
      }
    }
  } else if (NORM_A && (warp == 2 || warp == 3 || warp >= 8)) {
    // ------------------------------------------------------------------ A-operand producer (normalise + ReLU + split)
    // four warps (2, 3 and the two extra warps 8, 9 of the 320-thread NORM_A instantiation): one tile row per thread
    if constexpr (NORM_A && BK == 64) {
      const int tid = warp < 4 ? threadIdx.x - 64 : threadIdx.x - 192;  // 0..127
      // (mean, rstd) of the tile's (at most two) images, staged once per tile in the shared memory the statistics
      // epilogue would use (the two modes never meet in one GEMM): [2][K] float2, K <= 512
      float2* s_ms = s_stat;
      const int K = p.KB * BK;
      int it = 0;
      for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x) {
        int m0, n0, z;
        tile_coords(t, m0, n0, z);
        const int img_a = min(m0, p.M - 1) / p.a_np;
        gemm_norm_a_stage_stats(p, s_ms, tid, img_a, K);
        for (int kb = 0; kb < p.KB; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          uint8_t* st = smem + s * S::STAGE_BYTES;
          gemm_norm_a_row(p, st, s_ms, tid, m0, img_a, K, kb);
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the UMMA (async proxy) reads
          mbar_arrive(&a_ready_bar[s]);
        }
      }
    }
  } else if (warp >= 4) {
    int lt = 0;
    for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x, ++lt) {
      int m0, n0, z;
      tile_coords(t, m0, n0, z);
      const int a = lt % ACC;
      gemm_epilogue_tile<BN>(p, tmem_base + a * BN, s_bias + a * BN, s_stage, s_stat, m0, n0, z, [&]() {
        mbar_wait(&tmem_full_bar[a], (lt / ACC) & 1);
      });
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[a]);  // 128 arrivals: the accumulator may be overwritten
      if (p.stat_partial != nullptr) gemm_epilogue_stats_flush<BN>(p, s_stat, m0, n0);
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ACC * BN);
  }
}

template <int BN, int STAGES, int BK, bool NORM_A = false>
static int launch_impl(const GemmArgs& g, cudaStream_t stream) {
  using S = GemmSmem<BN, BK>;
  constexpr int SWZ = BK * 2;
  const bool split = g.nsplit == 3;
  CUtensorMap ta_hi, ta_lo, ta2_hi, ta2_lo, tb_hi, tb_lo;
  const int Kt = g.K1 + g.K2;
  if (NORM_A) {
    if (make_tmap_f32_3d(&ta_hi, g.a_f32, g.K1, g.M, g.batch, g.a_row_stride, g.a_batch_stride, GEMM_BM)) return 3;
  } else {
    if (make_tmap_f16_3d(&ta_hi, g.a_hi, g.K1, g.M, g.batch, g.a_row_stride, g.a_batch_stride, BK, GEMM_BM, SWZ)) return 3;
  }
  ta_lo = ta_hi;
  if (!NORM_A && split && make_tmap_f16_3d(&ta_lo, g.a_lo, g.K1, g.M, g.batch, g.a_row_stride, g.a_batch_stride, BK, GEMM_BM, SWZ)) return 3;
  ta2_hi = ta_hi;
  ta2_lo = ta_lo;
  if (g.K2 > 0) {
    if (make_tmap_f16_3d(&ta2_hi, g.a2_hi, g.K2, g.M, g.batch, g.a2_row_stride, g.a2_batch_stride, BK, GEMM_BM, SWZ)) return 3;
    ta2_lo = ta2_hi;
    if (split && make_tmap_f16_3d(&ta2_lo, g.a2_lo, g.K2, g.M, g.batch, g.a2_row_stride, g.a2_batch_stride, BK, GEMM_BM, SWZ)) return 3;
  }
  const int bb = g.b_batched ? g.batch : 1;
  if (make_tmap_f16_3d(&tb_hi, g.b_hi, Kt, g.N, bb, g.b_row_stride, g.b_batch_stride, BK, BN, SWZ)) return 3;
  tb_lo = tb_hi;
  if (split && make_tmap_f16_3d(&tb_lo, g.b_lo, Kt, g.N, bb, g.b_row_stride, g.b_batch_stride, BK, BN, SWZ)) return 3;

  GemmKernelParams p;
  p.M = g.M;
  p.N = g.N;
  p.KB1 = g.K1 / BK;
  p.KB = Kt / BK;
  p.b_batched = g.b_batched;
  p.nsplit = g.nsplit;
  p.out_mode = g.out_mode;
  p.alpha = g.alpha;
  p.bias = g.bias;
  p.out0 = g.out0;
  p.out1 = g.out1;
  p.res_hi = reinterpret_cast<const __half*>(g.res_hi);
  p.res_lo = reinterpret_cast<const __half*>(g.res_lo);
  p.out_row_stride = g.out_row_stride;
  p.out_batch_stride = g.out_batch_stride;
  p.stat_partial = reinterpret_cast<float2*>(g.stat_partial);
  p.stat_straddle = reinterpret_cast<float2*>(g.stat_straddle);
  p.stat_ns = g.stat_ns;
  p.stat_np = g.stat_np;
  p.a_stats = reinterpret_cast<const float2*>(g.a_stats);
  p.a_np = g.a_np;

  const size_t smem = STAGES * S::STAGE_BYTES + 1024 + 256 + 2 * BN * sizeof(float) + 4 * 32 * 144 +
                      (BN == 256 ? 4 * BN * sizeof(float2) : 0);
  auto kern = gemm_f16split_kernel<BN, STAGES, BK, NORM_A>;
  static DeviceOnce configured;
  if (configured.first()) {
    IMP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  p.tiles_n = (g.N + BN - 1) / BN;
  p.tiles_m = (g.M + GEMM_BM - 1) / GEMM_BM;
  p.tiles_total = p.tiles_n * p.tiles_m * g.batch;
  const int grid = p.tiles_total < num_sms() ? p.tiles_total : num_sms();
  kern<<<grid, NORM_A ? GEMM_THREADS + 64 : GEMM_THREADS, smem, stream>>>(ta_hi, ta_lo, ta2_hi, ta2_lo, tb_hi, tb_lo, p);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}


// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster (the two SMs of a TPC) compute one 256 x 256 output tile with
// M = 256 MMAs.  Each CTA stages ITS 128 rows of A and ITS 128 of the tile's 256 B rows; the tensor cores of both SMs read
// both halves of B, so a CTA fetches (128 + 128) x K operand rows per 128 x 256 outputs instead of (128 + 256) x K -- the
// single-CTA kernel is bound by exactly that L2 -> shared-memory traffic (3-product split: 4 planes per operand tile).
// The smaller stage (64 KB) also buys a third pipeline stage.  Roles as above; the even CTA ("leader") owns the full /
// tmem_empty barriers and issues every MMA, commits are multicast to both CTAs, each CTA drains its own 128 accumulator
// rows with the shared epilogue code.
static constexpr int GP_BN = 256, GP_BK = 64, GP_STAGES = 3;
struct GemmPairSmem {
  static constexpr int A_BYTES = GEMM_BM * GP_BK * 2;        // 16 KB: 128 rows of one A plane
  static constexpr int B_BYTES = (GP_BN / 2) * GP_BK * 2;    // 16 KB: this CTA's 128 rows of one B plane
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int SBO = 8 * GP_BK * 2;
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_f16split_pair_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                          const __grid_constant__ CUtensorMap tm_a2_hi, const __grid_constant__ CUtensorMap tm_a2_lo,
                          const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                          const GemmKernelParams p) {
  using S = GemmPairSmem;
  constexpr int BN = GP_BN, BK = GP_BK, STAGES = GP_STAGES, ACC = 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);  // used in the leader
  uint64_t* empty_bar = full_bar + STAGES;                                         // one per CTA (multicast commits)
  uint64_t* tmem_full_bar = empty_bar + STAGES;                                    // [ACC] one per CTA
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC;                                  // [ACC] used in the leader
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + ACC);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr_smem + 4) + 15) & ~uintptr_t(15));
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_bias + ACC * BN);
  float2* s_stat = reinterpret_cast<float2*>(s_stage + 4 * 32 * 144);

  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const bool split = p.nsplit == 3;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_b_hi);
    if (split) {
      tma_prefetch_desc(&tm_a_lo);
      tma_prefetch_desc(&tm_b_lo);
    }
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);   // one arrive.expect_tx per CTA of the pair
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 256);  // the epilogue threads of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_ptr_smem, ACC * BN);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // work item t -> (n tile fastest, then 256-row m tile, then batch); this CTA owns rows m0 .. m0+127 of the pair's tile
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  auto tile_coords = [&](int t, int& m0, int& n0, int& z) {
    const int per_z = p.tiles_n * p.tiles_m;
    z = t / per_z;
    const int r = t - z * per_z;
    m0 = (r / p.tiles_n) * (2 * GEMM_BM) + (int)rank * GEMM_BM;
    n0 = (r % p.tiles_n) * BN;
  };

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t tx = split ? S::STAGE_BYTES : (S::A_BYTES + S::B_BYTES);
      int it = 0;
      for (int t = pair_id; t < p.tiles_total; t += n_pairs) {
        int m0, n0, z;
        tile_coords(t, m0, n0, z);
        const int nb = n0 + (int)rank * (BN / 2);  // this CTA's half of the tile's B rows
        for (int kb = 0; kb < p.KB; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* st = smem + s * S::STAGE_BYTES;
          mbar_arrive_expect_tx_cluster(mapa_u32(smem_u32(&full_bar[s]), 0), tx);
          const bool seg2 = kb >= p.KB1;
          const int ka = (seg2 ? kb - p.KB1 : kb) * BK;
          tma_load_3d_pair(st, seg2 ? &tm_a2_hi : &tm_a_hi, &full_bar[s], ka, m0, z);
          tma_load_3d_pair(st + 2 * S::A_BYTES, &tm_b_hi, &full_bar[s], kb * BK, nb, p.b_batched ? z : 0);
          if (split) {
            tma_load_3d_pair(st + S::A_BYTES, seg2 ? &tm_a2_lo : &tm_a_lo, &full_bar[s], ka, m0, z);
            tma_load_3d_pair(st + 2 * S::A_BYTES + S::B_BYTES, &tm_b_lo, &full_bar[s], kb * BK, nb, p.b_batched ? z : 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc(FMT_F16, 2 * GEMM_BM, BN, 0, 0);
      int it = 0, lt = 0;
      for (int t = pair_id; t < p.tiles_total; t += n_pairs, ++lt) {
        const int a = lt % ACC;
        mbar_wait(&tmem_empty_bar[a], ((lt / ACC) & 1) ^ 1);  // both CTAs have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        for (int kb = 0; kb < p.KB; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + s * S::STAGE_BYTES);
          const uint32_t a_lo = a_hi + S::A_BYTES;
          const uint32_t b_hi = a_hi + 2 * S::A_BYTES;
          const uint32_t b_lo = b_hi + S::B_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint32_t off = k * 32;
            const uint64_t dah = make_smem_desc(a_hi + off, 16, S::SBO, 2u);
            const uint64_t dbh = make_smem_desc(b_hi + off, 16, S::SBO, 2u);
            umma_f16_ss_pair(d_tmem, dah, dbh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (split) {
              const uint64_t dal = make_smem_desc(a_lo + off, 16, S::SBO, 2u);
              const uint64_t dbl = make_smem_desc(b_lo + off, 16, S::SBO, 2u);
              umma_f16_ss_pair(d_tmem, dal, dbh, idesc, 1u);
              umma_f16_ss_pair(d_tmem, dah, dbl, idesc, 1u);
            }
          }
          umma_commit_pair(&empty_bar[s]);  // frees the slot in both CTAs when these MMAs retire
        }
        umma_commit_pair(&tmem_full_bar[a]);
      }
    }
  } else if (warp >= 4) {
    int lt = 0;
    for (int t = pair_id; t < p.tiles_total; t += n_pairs, ++lt) {
      int m0, n0, z;
      tile_coords(t, m0, n0, z);
      const int a = lt % ACC;
      gemm_epilogue_tile<BN>(p, tmem_base + a * BN, s_bias + a * BN, s_stage, s_stat, m0, n0, z, [&]() {
        mbar_wait(&tmem_full_bar[a], (lt / ACC) & 1);
      });
      tc_fence_before();
      mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[a]), 0));  // 2 x 128 arrivals on the leader's barrier
      if (p.stat_partial != nullptr) gemm_epilogue_stats_flush<BN>(p, s_stat, m0, n0);
    }
  }
  tc_fence_before();
  cluster_sync_all();  // the peer may still signal this CTA's barriers / read its shared memory until here
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, ACC * BN);
  }
}

static int launch_pair(const GemmArgs& g, cudaStream_t stream) {
  using S = GemmPairSmem;
  constexpr int BN = GP_BN, BK = GP_BK;
  const bool split = g.nsplit == 3;
  CUtensorMap ta_hi, ta_lo, ta2_hi, ta2_lo, tb_hi, tb_lo;
  const int Kt = g.K1 + g.K2;
  if (make_tmap_f16_3d(&ta_hi, g.a_hi, g.K1, g.M, g.batch, g.a_row_stride, g.a_batch_stride, BK, GEMM_BM, 128)) return 3;
  ta_lo = ta_hi;
  if (split && make_tmap_f16_3d(&ta_lo, g.a_lo, g.K1, g.M, g.batch, g.a_row_stride, g.a_batch_stride, BK, GEMM_BM, 128)) return 3;
  ta2_hi = ta_hi;
  ta2_lo = ta_lo;
  if (g.K2 > 0) {
    if (make_tmap_f16_3d(&ta2_hi, g.a2_hi, g.K2, g.M, g.batch, g.a2_row_stride, g.a2_batch_stride, BK, GEMM_BM, 128)) return 3;
    ta2_lo = ta2_hi;
    if (split && make_tmap_f16_3d(&ta2_lo, g.a2_lo, g.K2, g.M, g.batch, g.a2_row_stride, g.a2_batch_stride, BK, GEMM_BM, 128)) return 3;
  }
  const int bb = g.b_batched ? g.batch : 1;
  if (make_tmap_f16_3d(&tb_hi, g.b_hi, Kt, g.N, bb, g.b_row_stride, g.b_batch_stride, BK, BN / 2, 128)) return 3;
  tb_lo = tb_hi;
  if (split && make_tmap_f16_3d(&tb_lo, g.b_lo, Kt, g.N, bb, g.b_row_stride, g.b_batch_stride, BK, BN / 2, 128)) return 3;

  GemmKernelParams p;
  p.M = g.M;
  p.N = g.N;
  p.KB1 = g.K1 / BK;
  p.KB = Kt / BK;
  p.b_batched = g.b_batched;
  p.nsplit = g.nsplit;
  p.out_mode = g.out_mode;
  p.alpha = g.alpha;
  p.bias = g.bias;
  p.out0 = g.out0;
  p.out1 = g.out1;
  p.res_hi = reinterpret_cast<const __half*>(g.res_hi);
  p.res_lo = reinterpret_cast<const __half*>(g.res_lo);
  p.out_row_stride = g.out_row_stride;
  p.out_batch_stride = g.out_batch_stride;
  p.stat_partial = reinterpret_cast<float2*>(g.stat_partial);
  p.stat_straddle = reinterpret_cast<float2*>(g.stat_straddle);
  p.stat_ns = g.stat_ns;
  p.stat_np = g.stat_np;
  p.a_stats = reinterpret_cast<const float2*>(g.a_stats);
  p.a_np = g.a_np;
  p.tiles_n = (g.N + BN - 1) / BN;
  p.tiles_m = (g.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);  // 256-row tiles of a pair
  p.tiles_total = p.tiles_n * p.tiles_m * g.batch;

  const size_t smem = GP_STAGES * S::STAGE_BYTES + 1024 + 256 + 2 * BN * sizeof(float) + 4 * 32 * 144 + 4 * BN * sizeof(float2);
  static DeviceOnce configured;
  if (configured.first()) {
    IMP_CUDA_OK(cudaFuncSetAttribute(gemm_f16split_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int max_pairs = num_sms() / 2;
  const int pairs = p.tiles_total < max_pairs ? p.tiles_total : max_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  IMP_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_f16split_pair_kernel, ta_hi, ta_lo, ta2_hi, ta2_lo, tb_hi, tb_lo, p));
  return 0;
}

static int g_gemm_variant = -1;
void gemm_set_variant(int v) { g_gemm_variant = v; }

int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
  IMP_REQUIRE(g.nsplit == 1 || g.nsplit == 3, "gemm: nsplit must be 1 or 3");
  IMP_REQUIRE(g.K1 > 0 && g.K1 % 64 == 0 && g.K2 % 64 == 0, "gemm: K segments must be multiples of 64 (got %d, %d)", g.K1, g.K2);
  IMP_REQUIRE(g.M > 0 && g.N > 0 && g.batch > 0, "gemm: empty problem");
  IMP_REQUIRE(g.out_row_stride % 8 == 0, "gemm: output row stride must be a multiple of 8 elements");
  if (g.stat_partial != nullptr)
    IMP_REQUIRE(g.out_mode == IMP_GEMM_OUT_F32 && g.N > 128 && g.N % 32 == 0 && g.batch == 1 && g.stat_np >= GEMM_BM &&
                    g.stat_ns != nullptr && g.stat_straddle != nullptr && g.M % g.stat_np == 0,
                "gemm: instance-norm statistics need fp32 output, 128 < N (multiple of 32), batch 1 and images of >= 128 rows");
  if (g.a_f32 != nullptr) {
    IMP_REQUIRE(g.a_stats != nullptr && g.a_np >= GEMM_BM && g.M % g.a_np == 0 && g.K2 == 0 && g.K1 <= 512 && g.nsplit == 3 &&
                    g.N > 128 && g.batch == 1,
                "gemm: the normalising A path needs a_stats, images of >= 128 rows, one K segment <= 512, nsplit 3, N > 128, batch 1");
    // (a CTA-pair version of this path -- 64 KB stages, three of them -- measured slower: 0.376 vs 0.350 ms; removed)
    return launch_impl<256, 2, 64, true>(g, stream);
  }
  if (g_gemm_variant < 0) {
    const char* e = getenv("IMP_GEMM_VARIANT");
    g_gemm_variant = e ? atoi(e) : 0;
  }
  const int variant = g_gemm_variant;
  // CTA pairs (cta_group::2, 256 x 256 tiles) where they measured faster and the problem fills the 74 pairs at least twice:
  // the wide, tensor-bound projections (QKV N = 768: 0.289 -> 0.273 ms, MLP0 N = 512: 0.346 -> 0.316 ms at 256 k rows; bit-identical
  // results).  N = 256 GEMMs (HBM-bound) and the batched score GEMM measure the same either way and stay on the single-CTA
  // kernel, as do small problems (one pair per call).  IMP_GEMM_VARIANT: 2 = pairs wherever N > 128, 3 = never.
  if (g.N > 128 && variant == 2) return launch_pair(g, stream);
  if (variant == 0 && g.N >= 512 && !g.b_batched) {
    const long long pair_tiles = (long long)((g.N + GP_BN - 1) / GP_BN) * ((g.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM)) * g.batch;
    if (pair_tiles >= 2LL * (num_sms() / 2)) return launch_pair(g, stream);
  }
  // two 96 KB stages (BK = 64) and four 48 KB stages (BK = 32) measure the same on B200 (tools/gemm_probe.py)
  if (g.N > 128) return variant != 1 ? launch_impl<256, 2, 64>(g, stream) : launch_impl<256, 4, 32>(g, stream);
  return launch_impl<128, 3, 64>(g, stream);
}

}  // namespace imp
