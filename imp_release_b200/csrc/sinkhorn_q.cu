// Sinkhorn for big batches (nets/layers.py:27-46, same recurrence as sinkhorn.cu), "column-split" streaming kernels.
//
// At 64 pairs x 2001^2 the 19 iteration sweeps are pure HBM streaming, so they get faster only by moving fewer bytes and
// by keeping the SM-side cost per element far below the HBM rate:
//  * storage: the init pass writes p = softmax_rows(pad(dist)) as fp32, as a 24-bit copy (top 16 bits of the fp32 word +
//    one byte of mantissa extension, planar) or as p * 2^14 in IEEE fp16; the final pass re-derives p in fp32 from dist
//    and the saved row statistics (max, 1/sum), so the rounding of the copy only enters through the scaling vectors u, v
//    (measured deviations: DESIGN.md section 2).  All arithmetic on decoded values is fp32.
//  * work split: a CTA owns a block of rows; its 8 consumer warps split the COLUMNS, every lane owning 8 fixed columns.
//    v_j, the column-sum accumulators and (final pass) the column arg-max trackers of those columns live in registers for
//    the whole CTA, so shared memory is touched exactly once per element (the bulk-TMA landing zone).  Rows are handled
//    T at a time: per-row partial sums are reduced with a transposed shuffle tree (9 shuffles for 8 rows), exchanged
//    through a few floats of shared memory with ONE named barrier per T rows, and the slot is handed back to the producer
//    warp as soon as the rows sit in registers.
//  * one sweep per iteration (u_i, then p_ij u_i folded into the column sums the next launch turns into v), and the final
//    pass also produces the column arg-max (packed atomicMax per owned column), so a scoring is 1 + 19 + 1 sweeps.
#include <cuda_fp16.h>
#include <stdlib.h>

#include <type_traits>

#include "common.h"
#include "sinkhorn_common.cuh"

namespace imp {

static constexpr int SKQ_CW = 8;                          // consumer warps (column owners)
static constexpr int SKQ_THREADS = (SKQ_CW + 1) * 32;     // + one producer warp
static constexpr int SKQ_SMEM_BUDGET = 113 * 1024;        // two CTAs per SM
static constexpr int SKQ_FIXED_SMEM = 4096;               // partial-sum exchange + barriers
static constexpr float SKQ_F16_SCALE = 16384.f;           // p <= 1 -> <= 2^14; fp16 normals reach down to p = 3.7e-9
static constexpr float SKQ_F16_INV = 1.f / 16384.f;

enum { QF32 = IMP_SK_STORE_F32, QF16 = IMP_SK_STORE_F16, QF24 = IMP_SK_STORE_F24 };
template <int FMT>
struct QFmt {
  static constexpr int BPE = FMT == QF32 ? 4 : (FMT == QF16 ? 2 : 3);  // bytes per element
  static constexpr int RAW = FMT == QF32 ? 8 : (FMT == QF16 ? 4 : 6);  // 32-bit registers per 8 elements
};

struct SkqParams {
  const float* dist;
  long long dist_bs;
  int ldd;
  const float* bin_score;
  unsigned char* Q;  // per matrix [Rmax][ldq] x {fp32 | fp16 | 16-bit plane followed by an 8-bit plane}
  long long q_bs;    // bytes
  int ldq, Rmax;
  float* P;
  long long p_bs;
  int ldp;
  float* row_m;    // [batch][Rmax] row max of the padded logits
  float* row_inv;  // [batch][Rmax] 1 / sum exp(x - max)
  float* u;
  const float* col_prev;
  float* col_acc;
  float* col_zero;
  int ldc;
  float* row_max;
  int* row_arg;
  float* row_mass;
  float* col_mass;
  unsigned long long* col_key;
  const int *n0s, *n1s;
  int N0max, N1max;
  int rows_per_cta, blocks_per_mat, n_items, nslots, slot_bytes, row_bytes;
  int do_iter, write_scores;
};

// mbarrier wait whose polls are suspended in hardware for up to ~20 us at a time: the plain try_wait loop of ptx.cuh
// re-polls every few hundred cycles, and in these memory-bound kernels the polls of the waiting warps were 14 % of all
// issued instructions (ncu), competing with the warps that have work
__device__ __forceinline__ void skq_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1, 20000;\n\t"
      "@!P bra WAIT_%=;\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// ring position of the running batch counter, kept incrementally (n % nslots and n / nslots by a run-time nslots cost
// two integer divisions per batch)
__device__ __forceinline__ void skq_ring_next(int& slot, uint32_t& phase, int nslots) {
  if (++slot == nslots) {
    slot = 0;
    phase ^= 1u;
  }
}
__device__ __forceinline__ void skq_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(SKQ_CW * 32) : "memory"); }

// Hand a ring slot back to the producer.  The arrive must not overtake the shared-memory loads that filled the consumer's
// registers: an LDS that is issued but not yet performed is NOT ordered before a later mbarrier.arrive by the hardware
// (measured: a few rows per launch read the refilled slot), so the arrive is made data-dependent on `dep`, a value that
// every lane's loads of the slot feed through a warp shuffle reduction (bits 0xffffffff = a NaN pattern the reductions
// never produce, so the predicate is always true, but ptxas cannot know).
__device__ __forceinline__ void skq_release(uint64_t* bar, int lane, float dep) {
  if (lane == 0)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %1, 0xffffffff;\n\t"
        "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(__float_as_uint(dep))
        : "memory");
}

// logits of one float4 group of the padded matrix: dustbin column / row = bin, -FLT_MAX beyond C
__device__ __forceinline__ float4 skq_logits(const float* srow, int c0, int C, bool bin_row, float bin) {
  if (!bin_row && c0 + 3 < C - 1) return *reinterpret_cast<const float4*>(srow + c0);  // interior group
  float4 t = make_float4(bin, bin, bin, bin);
  if (!bin_row && c0 < C - 1) t = *reinterpret_cast<const float4*>(srow + c0);
  t.x = (c0 + 0 < C) ? ((bin_row || c0 + 0 == C - 1) ? bin : t.x) : -FLT_MAX;
  t.y = (c0 + 1 < C) ? ((bin_row || c0 + 1 == C - 1) ? bin : t.y) : -FLT_MAX;
  t.z = (c0 + 2 < C) ? ((bin_row || c0 + 2 == C - 1) ? bin : t.z) : -FLT_MAX;
  t.w = (c0 + 3 < C) ? ((bin_row || c0 + 3 == C - 1) ? bin : t.w) : -FLT_MAX;
  return t;
}
// the same with the per-lane case analysis hoisted out of the row loop: `interior` (all four columns < C-1) is a property
// of the lane's group and the work item; only the one boundary group and the groups beyond C take the slow path
__device__ __forceinline__ float4 skq_logits_fast(const float* srow, int c0, int C, bool interior, bool bin_row, float bin) {
  if (interior && !bin_row) return *reinterpret_cast<const float4*>(srow + c0);
  return skq_logits(srow, c0, C, bin_row, bin);
}
// e = exp(x - max) in two instructions, 2^(x log2e - max log2e): the init and final passes are issue-bound, not HBM-bound,
// with the 8-instruction sk_exp.  Relative error ~ 4e-8 |t| (t = the exponent; the rounding of the FMA) + 2 ulp of
// MUFU.EX2, i.e. < 7e-7 wherever the probability exceeds 1e-3 of the row maximum; both passes use this one expression
// with the stored row max, so they derive identical bits.  -FLT_MAX (pad) overflows to -inf -> exactly 0.
__device__ __forceinline__ float skq_exp(float x, float max_log2e) {
  return fast_exp2(__fmaf_rn(x, 1.4426950408889634f, -max_log2e));
}
__device__ __forceinline__ float skq_max_log2e(float m) { return __fmul_rn(m, 1.4426950408889634f); }

// 24-bit encoding of a non-negative fp32: the top 16 bits verbatim, the low 16 bits L rounded to the nearest multiple of
// 257 (q = (L + 128) / 257 in 0..255, no carry since 255 * 257 = 65535) so that a single byte-permute rebuilds the word
// as [b3 b2 q q]; |error| <= 128.5 ulp(fp32) = 1.5e-5 relative.
__device__ __forceinline__ void skq_enc24(float x, uint32_t& hi, uint32_t& lo) {
  const uint32_t bits = __float_as_uint(x);
  hi = bits >> 16;
  lo = ((bits & 0xFFFFu) * 65281u + 128u * 65281u) >> 24;  // == ((bits & 0xFFFF) + 128) / 257 for all 16-bit inputs
}

template <int FMT>
__device__ __forceinline__ void skq_store4(unsigned char* qrow, unsigned char* qlo, int c0, float4 t) {
  if (FMT == QF32) {
    *reinterpret_cast<float4*>(qrow + 4 * c0) = t;
  } else if (FMT == QF16) {
    uint2 e;
    e.x = pack_half2(t.x * SKQ_F16_SCALE, t.y * SKQ_F16_SCALE);
    e.y = pack_half2(t.z * SKQ_F16_SCALE, t.w * SKQ_F16_SCALE);
    *reinterpret_cast<uint2*>(qrow + 2 * c0) = e;
  } else {
    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
    skq_enc24(t.x, h0, l0);
    skq_enc24(t.y, h1, l1);
    skq_enc24(t.z, h2, l2);
    skq_enc24(t.w, h3, l3);
    *reinterpret_cast<uint2*>(qrow + 2 * c0) = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
    *reinterpret_cast<uint32_t*>(qlo + c0) = l0 | (l1 << 8) | (l2 << 16) | (l3 << 24);
  }
}

// eight consecutive elements of a staged row: raw registers <- shared memory, fp32 <- raw registers
template <int FMT>
__device__ __forceinline__ void skq_load8(const unsigned char* srow, int lo_off, int c0, uint32_t (&r)[QFmt<FMT>::RAW]) {
  if constexpr (FMT == QF32) {
    const uint4 a = *reinterpret_cast<const uint4*>(srow + 4 * c0);
    const uint4 b = *reinterpret_cast<const uint4*>(srow + 4 * c0 + 16);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w;
    r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
  } else {
    const uint4 h = *reinterpret_cast<const uint4*>(srow + 2 * c0);
    r[0] = h.x; r[1] = h.y; r[2] = h.z; r[3] = h.w;
    if constexpr (FMT == QF24) {
      const uint2 l = *reinterpret_cast<const uint2*>(srow + lo_off + c0);
      r[4] = l.x;
      r[5] = l.y;
    }
  }
}
template <int FMT>
__device__ __forceinline__ void skq_decode8(const uint32_t (&r)[QFmt<FMT>::RAW], float (&f)[8]) {
  if constexpr (FMT == QF32) {
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] = __uint_as_float(r[q]);
  } else if constexpr (FMT == QF16) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r[q]));
      f[2 * q] = a.x;
      f[2 * q + 1] = a.y;
    }
  } else {
    f[0] = __uint_as_float(__byte_perm(r[0], r[4], 0x1044));
    f[1] = __uint_as_float(__byte_perm(r[0], r[4], 0x3255));
    f[2] = __uint_as_float(__byte_perm(r[1], r[4], 0x1066));
    f[3] = __uint_as_float(__byte_perm(r[1], r[4], 0x3277));
    f[4] = __uint_as_float(__byte_perm(r[2], r[5], 0x1044));
    f[5] = __uint_as_float(__byte_perm(r[2], r[5], 0x3255));
    f[6] = __uint_as_float(__byte_perm(r[3], r[5], 0x1066));
    f[7] = __uint_as_float(__byte_perm(r[3], r[5], 0x3277));
  }
}

// Transposed warp reduction of T per-row values: after log2(T) "keep one half, send the other" steps every lane holds
// ONE row's partial, the remaining steps are plain butterflies.  Returns the warp total of row skq_rid<T>(lane) (the same
// in the 32/T lanes that share that row id).  9 shuffles for T = 8 instead of 40.
template <int T>
__device__ __forceinline__ int skq_rid(int lane) {
  return T == 8 ? (((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1))
                : (T == 4 ? (((lane >> 4) & 1) * 2 + ((lane >> 3) & 1)) : ((lane >> 4) & 1));
}
template <int T, typename Op>
__device__ __forceinline__ float skq_tr_reduce(float (&a)[T], int lane, Op op) {
  static_assert(T == 2 || T == 4 || T == 8, "T must be 2, 4 or 8");
  int h = 16;
#pragma unroll
  for (int n = T / 2; n >= 1; n >>= 1, h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = up ? a[i] : a[i + n];
      const float keep = up ? a[i + n] : a[i];
      a[i] = op(keep, __shfl_xor_sync(0xffffffffu, send, h));
    }
  }
  float r = a[0];
  for (; h >= 1; h >>= 1) r = op(r, __shfl_xor_sync(0xffffffffu, r, h));
  return r;
}

// Work item = one block of rows_per_cta rows of one matrix.  The kernels are persistent: CTA k walks items k, k + grid,
// k + 2 grid, ... with one running batch counter, so the bulk-copy ring keeps streaming across item boundaries.
struct SkqCta {
  int b, row0, nrows, R, C, Cq;
};
__device__ __forceinline__ bool skq_item(const SkqParams& p, int item, SkqCta& c) {
  c.b = item / p.blocks_per_mat;
  const int bx = item - c.b * p.blocks_per_mat;
  const SkDims d = sk_dims(p.n0s, p.n1s, c.b, p.N0max, p.N1max);
  c.R = d.R;
  c.C = d.C;
  c.Cq = (d.C + 15) & ~15;
  c.row0 = bx * p.rows_per_cta;
  if (c.row0 >= d.R) return false;
  c.nrows = min(p.rows_per_cta, d.R - c.row0);
  return true;
}

struct SkqSmem {
  unsigned char* ring;
  float* part;  // exchange area, 512 floats
  uint64_t *full_bar, *empty_bar;
};
__device__ __forceinline__ SkqSmem skq_smem_setup(unsigned char* base, const SkqParams& p) {
  SkqSmem s;
  s.ring = base;
  s.part = reinterpret_cast<float*>(base + (size_t)p.nslots * p.slot_bytes);
  s.full_bar = reinterpret_cast<uint64_t*>(s.part + 512);
  s.empty_bar = s.full_bar + p.nslots;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nslots; ++i) {
      mbar_init(&s.full_bar[i], 1);
      mbar_init(&s.empty_bar[i], SKQ_CW);
    }
    fence_barrier_init();
  }
  __syncthreads();
  return s;
}

// producer warp for the passes that stream dist rows (init, final): T rows per slot
template <int T>
__device__ __forceinline__ void skq_produce_dist(const SkqParams& p, const SkqSmem& s) {
  if (lane_id() != 0) return;
  int n = 0, slot = 0;  // running batch counter of this CTA and its ring position
  uint32_t phase = 0;
  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
    SkqCta c;
    if (!skq_item(p, item, c)) continue;
    const float* src = p.dist + c.b * p.dist_bs;
    const uint32_t bytes = (uint32_t)(((c.C - 1 + 3) & ~3) * 4);
    const int nbatch = (c.nrows + T - 1) / T;
    for (int j = 0; j < nbatch; ++j, ++n, skq_ring_next(slot, phase, p.nslots)) {
      skq_wait(&s.empty_bar[slot], phase ^ 1u);
      const int i0 = c.row0 + j * T;
      const int nb = min(T, c.nrows - j * T);
      int ncopy = 0;
      for (int t = 0; t < nb; ++t) ncopy += (i0 + t != c.R - 1 && bytes != 0) ? 1 : 0;  // the dustbin row has no source
      if (ncopy == 0) {
        mbar_arrive(&s.full_bar[slot]);
        continue;
      }
      mbar_arrive_expect_tx(&s.full_bar[slot], bytes * ncopy);
      unsigned char* dst = s.ring + (size_t)slot * p.slot_bytes;
      for (int t = 0; t < nb; ++t)
        if (i0 + t != c.R - 1 && bytes != 0)
          bulk_copy_g2s(dst + (size_t)t * p.row_bytes, src + (long long)(i0 + t) * p.ldd, bytes, &s.full_bar[slot]);
    }
  }
}

// producer warp for a pass that streams the stored fp32 copy (final pass with fp32 storage): T rows per slot
template <int T>
__device__ __forceinline__ void skq_produce_q32(const SkqParams& p, const SkqSmem& s) {
  if (lane_id() != 0) return;
  int n = 0, slot = 0;
  uint32_t phase = 0;
  const size_t stride = (size_t)p.ldq * 4;
  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
    SkqCta c;
    if (!skq_item(p, item, c)) continue;
    const unsigned char* q0 = p.Q + c.b * p.q_bs;
    const uint32_t bytes = (uint32_t)c.Cq * 4;
    const int nbatch = (c.nrows + T - 1) / T;
    for (int j = 0; j < nbatch; ++j, ++n, skq_ring_next(slot, phase, p.nslots)) {
      skq_wait(&s.empty_bar[slot], phase ^ 1u);
      const int i0 = c.row0 + j * T;
      const int nb = min(T, c.nrows - j * T);
      mbar_arrive_expect_tx(&s.full_bar[slot], bytes * nb);
      unsigned char* dst = s.ring + (size_t)slot * p.slot_bytes;
      for (int t = 0; t < nb; ++t) bulk_copy_g2s(dst + (size_t)t * p.row_bytes, q0 + (size_t)(i0 + t) * stride, bytes, &s.full_bar[slot]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// init: p = softmax_rows(pad(dist)) -> stored copy + row statistics (+ first half-iteration: u with v = 1 and the
// column sums with that u, from the exact fp32 p).  Lanes own NG = 2 * NVW float4 column groups.
template <int FMT, int NVW, int T>
__global__ void __launch_bounds__(SKQ_THREADS, 2) skq_init_kernel(const SkqParams p) {
  extern __shared__ __align__(16) unsigned char skq_smem[];
  constexpr int NG = 2 * NVW;
  const SkqSmem s = skq_smem_setup(skq_smem, p);
  const int warp = threadIdx.x >> 5, lane = lane_id();
  if (warp == SKQ_CW) {
    skq_produce_dist<T>(p, s);
    return;
  }
  float* s_max = s.part;              // [T][CW]
  float* s_sum = s.part + T * SKQ_CW; // [T][CW]
  const float bin = *p.bin_score;
  const int rid = skq_rid<T>(lane);
  int c0s[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) c0s[k] = 4 * ((warp * NG + k) * 32 + lane);
  int n = 0, slot = 0;  // running batch counter (matches the producer's) and its ring position
  uint32_t phase = 0;
  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
  SkqCta c;
  if (!skq_item(p, item, c)) continue;
  const int b = c.b;
  if (c.row0 == 0 && p.col_zero != nullptr)
    for (int j = threadIdx.x; j < p.ldc; j += SKQ_CW * 32) p.col_zero[(long long)b * p.ldc + j] = 0.f;
  float4 acc[NG];
  bool interior[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    interior[k] = c0s[k] + 3 < c.C - 1;
  }

  const int nbatch = (c.nrows + T - 1) / T;
  for (int jb = 0; jb < nbatch; ++jb, ++n, skq_ring_next(slot, phase, p.nslots)) {
    const int nb = min(T, c.nrows - jb * T);
    const int i0 = c.row0 + jb * T;
    skq_wait(&s.full_bar[slot], phase);
    const unsigned char* sb = s.ring + (size_t)slot * p.slot_bytes;
    // pass 1: padded logits into registers, per-row max
    float4 x[T][NG];
    float red[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      red[t] = -FLT_MAX;
      if (t < nb) {
        const bool bin_row = (i0 + t == c.R - 1);
        const float* srow = reinterpret_cast<const float*>(sb + (size_t)t * p.row_bytes);
#pragma unroll
        for (int k = 0; k < NG; ++k) {  // groups beyond C come back as -FLT_MAX: exp -> 0, nothing to guard below
          x[t][k] = skq_logits_fast(srow, c0s[k], c.C, interior[k], bin_row, bin);
          red[t] = fmaxf(red[t], fmaxf(fmaxf(x[t][k].x, x[t][k].y), fmaxf(x[t][k].z, x[t][k].w)));
        }
      }
    }
    {
      const float m = skq_tr_reduce<T>(red, lane, [](float a, float b2) { return fmaxf(a, b2); });
      skq_release(&s.empty_bar[slot], lane, m);  // every lane's logits sit in registers (they fed m)
      if ((lane & (32 / T - 1)) == 0) s_max[rid * SKQ_CW + warp] = m;
    }
    skq_consumer_sync();
    float my_m = -FLT_MAX;
    if (lane < T) {
      const float4 a = *reinterpret_cast<const float4*>(s_max + lane * SKQ_CW);
      const float4 d = *reinterpret_cast<const float4*>(s_max + lane * SKQ_CW + 4);
      my_m = fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(d.x, d.y), fmaxf(d.z, d.w)));
    }
    // pass 2: e = exp(x - max), per-row sum
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float mL = skq_max_log2e(__shfl_sync(0xffffffffu, my_m, t));
      red[t] = 0.f;
      if (t < nb) {
#pragma unroll
        for (int k = 0; k < NG; ++k) {
          x[t][k].x = skq_exp(x[t][k].x, mL);
          x[t][k].y = skq_exp(x[t][k].y, mL);
          x[t][k].z = skq_exp(x[t][k].z, mL);
          x[t][k].w = skq_exp(x[t][k].w, mL);
          red[t] += (x[t][k].x + x[t][k].y) + (x[t][k].z + x[t][k].w);
        }
      }
    }
    {
      const float sm = skq_tr_reduce<T>(red, lane, [](float a, float b2) { return a + b2; });
      if ((lane & (32 / T - 1)) == 0) s_sum[rid * SKQ_CW + warp] = sm;
    }
    skq_consumer_sync();
    float my_inv = 0.f, my_u = 0.f;
    if (lane < T) {
      const float4 a = *reinterpret_cast<const float4*>(s_sum + lane * SKQ_CW);
      const float4 d = *reinterpret_cast<const float4*>(s_sum + lane * SKQ_CW + 4);
      const float sum = ((a.x + a.y) + (a.z + a.w)) + ((d.x + d.y) + (d.z + d.w));
      my_inv = 1.0f / sum;  // one division per row; p = e * (1/sum) differs from e / sum by <= 1 ulp
      // first half-iteration with v = 1: sum_j p_ij = sum * inv_sum to ~1e-7 relative
      my_u = p.do_iter ? ((i0 + lane == c.R - 1) ? (float)c.R : 1.f) / (sum * my_inv + SK_EPS) : 0.f;
      if (warp == 0 && lane < nb) {
        const long long o = (long long)b * p.Rmax + i0 + lane;
        if (p.do_iter) p.u[o] = my_u;
        p.row_m[o] = my_m;
        p.row_inv[o] = my_inv;
      }
    }
    // pass 3: normalise, encode, store; fold p * u into the column accumulators
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float inv = __shfl_sync(0xffffffffu, my_inv, t);
      const float ui = __shfl_sync(0xffffffffu, my_u, t);
      if (t < nb) {
        unsigned char* qrow = p.Q + b * p.q_bs + (size_t)(i0 + t) * p.ldq * (FMT == QF24 ? 2 : QFmt<FMT>::BPE);
        unsigned char* qlo = p.Q + b * p.q_bs + (size_t)p.Rmax * p.ldq * 2 + (size_t)(i0 + t) * p.ldq;
#pragma unroll
        for (int k = 0; k < NG; ++k) {
          float4 e = x[t][k];  // exactly 0 in the pad columns [C, Cq)
          e.x = __fmul_rn(e.x, inv);
          e.y = __fmul_rn(e.y, inv);
          e.z = __fmul_rn(e.z, inv);
          e.w = __fmul_rn(e.w, inv);
          if (c0s[k] < c.Cq) skq_store4<FMT>(qrow, qlo, c0s[k], e);
          acc[k].x = fmaf(e.x, ui, acc[k].x);
          acc[k].y = fmaf(e.y, ui, acc[k].y);
          acc[k].z = fmaf(e.z, ui, acc[k].z);
          acc[k].w = fmaf(e.w, ui, acc[k].w);
        }
      }
    }
  }
  if (p.do_iter) {
    float* ca = p.col_acc + (long long)b * p.ldc;
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      if (c0s[k] + 0 < c.C) atomicAdd(ca + c0s[k] + 0, acc[k].x);
      if (c0s[k] + 1 < c.C) atomicAdd(ca + c0s[k] + 1, acc[k].y);
      if (c0s[k] + 2 < c.C) atomicAdd(ca + c0s[k] + 2, acc[k].z);
      if (c0s[k] + 3 < c.C) atomicAdd(ca + c0s[k] + 3, acc[k].w);
    }
  }
  }  // items
}

// ---------------------------------------------------------------------------------------------------------------
// one Sinkhorn iteration in a single sweep over the stored copy: u_i = r_i / (sum_j p_ij v_j + eps), then p_ij u_i
// folded into the column sums the next launch turns into v.  Lanes own NVW groups of 8 columns.
template <int FMT, int NVW, int T>
__global__ void __launch_bounds__(SKQ_THREADS, 2) skq_iter_kernel(const SkqParams p) {
  extern __shared__ __align__(16) unsigned char skq_smem[];
  constexpr int RAW = QFmt<FMT>::RAW;
  const SkqSmem s = skq_smem_setup(skq_smem, p);
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int lo_off = p.ldq * 2;

  if (warp == SKQ_CW) {
    if (lane == 0) {
      const uint32_t bpe_main = FMT == QF24 ? 2 : QFmt<FMT>::BPE;
      const size_t main_stride = (size_t)p.ldq * bpe_main;
      int n = 0, slot = 0;  // running batch counter of this CTA and its ring position
  uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        SkqCta c;
        if (!skq_item(p, item, c)) continue;
        const unsigned char* q0 = p.Q + c.b * p.q_bs;
        const unsigned char* qlo = q0 + (size_t)p.Rmax * p.ldq * 2;
        const uint32_t main_bytes = (uint32_t)c.Cq * bpe_main, lo_bytes = (uint32_t)c.Cq;
        const int nbatch = (c.nrows + T - 1) / T;
        for (int jb = 0; jb < nbatch; ++jb, ++n, skq_ring_next(slot, phase, p.nslots)) {
              skq_wait(&s.empty_bar[slot], phase ^ 1u);
          const int i0 = c.row0 + jb * T;
          const int nb = min(T, c.nrows - jb * T);
          mbar_arrive_expect_tx(&s.full_bar[slot], (FMT == QF24 ? main_bytes + lo_bytes : main_bytes) * nb);
          unsigned char* dst = s.ring + (size_t)slot * p.slot_bytes;
          for (int t = 0; t < nb; ++t) {
            bulk_copy_g2s(dst + (size_t)t * p.row_bytes, q0 + (size_t)(i0 + t) * main_stride, main_bytes, &s.full_bar[slot]);
            if (FMT == QF24)
              bulk_copy_g2s(dst + (size_t)t * p.row_bytes + lo_off, qlo + (size_t)(i0 + t) * p.ldq, lo_bytes, &s.full_bar[slot]);
          }
        }
      }
    }
    return;
  }

  const float vscale = (FMT == QF16) ? SKQ_F16_INV : 1.f;
  const int rid = skq_rid<T>(lane);
  int c0s[NVW];
#pragma unroll
  for (int k = 0; k < NVW; ++k) c0s[k] = 8 * ((warp * NVW + k) * 32 + lane);
  int n = 0, slot = 0;  // running batch counter (matches the producer's) and its ring position
  uint32_t phase = 0;
  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
  SkqCta c;
  if (!skq_item(p, item, c)) continue;
  const int b = c.b;
  if (c.row0 == 0 && p.col_zero != nullptr)
    for (int j = threadIdx.x; j < p.ldc; j += SKQ_CW * 32) p.col_zero[(long long)b * p.ldc + j] = 0.f;
  // v_j of the owned columns (pre-scaled for the fp16 copy) and their column-sum accumulators.  Lanes whose columns lie
  // beyond the matrix read column group 0 with v = 0 instead of branching around every load (they flush nothing).
  float v[NVW][8], acc[NVW][8];
  int c0e[NVW];
#pragma unroll
  for (int k = 0; k < NVW; ++k) {
    c0e[k] = c0s[k] < c.Cq ? c0s[k] : 0;
    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
    if (c0s[k] < c.C) va = v_from_colsum(p.col_prev + (long long)b * p.ldc, c0s[k], c.C);  // C <= ldc, both multiples of 4 apart
    if (c0s[k] + 4 < c.C) vb = v_from_colsum(p.col_prev + (long long)b * p.ldc, c0s[k] + 4, c.C);
    v[k][0] = va.x * vscale; v[k][1] = va.y * vscale; v[k][2] = va.z * vscale; v[k][3] = va.w * vscale;
    v[k][4] = vb.x * vscale; v[k][5] = vb.y * vscale; v[k][6] = vb.z * vscale; v[k][7] = vb.w * vscale;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[k][q] = 0.f;
  }

  const int nbatch = (c.nrows + T - 1) / T;
  for (int jb = 0; jb < nbatch; ++jb, ++n, skq_ring_next(slot, phase, p.nslots)) {
    const int nb = min(T, c.nrows - jb * T);
    const int i0 = c.row0 + jb * T;
    skq_wait(&s.full_bar[slot], phase);
    const unsigned char* sb = s.ring + (size_t)slot * p.slot_bytes;
    float* part = s.part + (n & 1) * (T * SKQ_CW);
    // FULL = all T rows of the batch exist (every batch but the last of an item): no per-row guards
    auto batch = [&](auto full) {
      constexpr bool FULL = decltype(full)::value;
      uint32_t raw[T][NVW][RAW];
      float red[T];
#pragma unroll
      for (int t = 0; t < T; ++t)
        if (FULL || t < nb) {
#pragma unroll
          for (int k = 0; k < NVW; ++k) skq_load8<FMT>(sb + (size_t)t * p.row_bytes, lo_off, c0e[k], raw[t][k]);
        }
#pragma unroll
      for (int t = 0; t < T; ++t) {
        red[t] = 0.f;
        if (FULL || t < nb) {
          float r0 = 0.f, r1 = 0.f;
#pragma unroll
          for (int k = 0; k < NVW; ++k) {
            float f[8];
            skq_decode8<FMT>(raw[t][k], f);
            r0 += (f[0] * v[k][0] + f[1] * v[k][1]) + (f[2] * v[k][2] + f[3] * v[k][3]);
            r1 += (f[4] * v[k][4] + f[5] * v[k][5]) + (f[6] * v[k][6] + f[7] * v[k][7]);
          }
          red[t] = r0 + r1;
        }
      }
      {
        const float tot = skq_tr_reduce<T>(red, lane, [](float a, float b2) { return a + b2; });
        skq_release(&s.empty_bar[slot], lane, tot);  // every lane's rows sit in registers (they fed tot)
        if ((lane & (32 / T - 1)) == 0) part[rid * SKQ_CW + warp] = tot;
      }
      skq_consumer_sync();
      float my_u = 0.f;
      if (lane < T) {
        const float4 a = *reinterpret_cast<const float4*>(part + lane * SKQ_CW);
        const float4 d = *reinterpret_cast<const float4*>(part + lane * SKQ_CW + 4);
        const float rs = ((a.x + a.y) + (a.z + a.w)) + ((d.x + d.y) + (d.z + d.w));
        my_u = ((i0 + lane == c.R - 1) ? (float)c.R : 1.f) / (rs + SK_EPS);
        if (warp == 0 && lane < nb) p.u[(long long)b * p.Rmax + i0 + lane] = my_u;
      }
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const float uis = __shfl_sync(0xffffffffu, my_u, t) * vscale;
        if (FULL || t < nb) {
#pragma unroll
          for (int k = 0; k < NVW; ++k) {
            float f[8];
            skq_decode8<FMT>(raw[t][k], f);
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[k][q] = fmaf(f[q], uis, acc[k][q]);
          }
        }
      }
    };
    batch(std::false_type{});
  }
  // every column has exactly one owner lane in the CTA: straight to the global accumulators
  float* ca = p.col_acc + (long long)b * p.ldc;
#pragma unroll
  for (int k = 0; k < NVW; ++k)
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (c0s[k] + q < c.C) atomicAdd(ca + c0s[k] + q, acc[k][q]);
  }  // items
}

// ---------------------------------------------------------------------------------------------------------------
// final: out = (p u) v with p re-derived in fp32 from dist and the row statistics; row arg-max / masses over the
// non-dustbin block, column arg-max of the owned columns (packed atomicMax, lowest row wins ties); scores written to P
// only when the caller wants the matrix
struct SkqBest {
  float v;
  int j;
  float m;
};
__device__ __forceinline__ SkqBest skq_best_merge(SkqBest a, SkqBest o) {
  SkqBest r;
  const bool take = o.v > a.v || (o.v == a.v && o.j < a.j);  // lowest column wins ties
  r.v = take ? o.v : a.v;
  r.j = take ? o.j : a.j;
  r.m = a.m + o.m;
  return r;
}
__device__ __forceinline__ SkqBest skq_best_shfl(SkqBest a, int h) {
  SkqBest r;
  r.v = __shfl_xor_sync(0xffffffffu, a.v, h);
  r.j = __shfl_xor_sync(0xffffffffu, a.j, h);
  r.m = __shfl_xor_sync(0xffffffffu, a.m, h);
  return r;
}
template <int T>
__device__ __forceinline__ SkqBest skq_tr_reduce_best(SkqBest (&a)[T], int lane) {
  int h = 16;
#pragma unroll
  for (int n = T / 2; n >= 1; n >>= 1, h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const SkqBest send = up ? a[i] : a[i + n];
      const SkqBest keep = up ? a[i + n] : a[i];
      a[i] = skq_best_merge(keep, skq_best_shfl(send, h));
    }
  }
  SkqBest r = a[0];
  for (; h >= 1; h >>= 1) r = skq_best_merge(r, skq_best_shfl(r, h));
  return r;
}

// FROM_Q (fp32 storage): the stored copy IS the exact p = e * (1 / sum) the init pass derived, so the final pass streams it
// instead of dist and skips the re-derivation (exp, dustbin / pad case analysis): same bits, ~1/3 fewer instructions in a
// pass that is instruction-bound (3 TB/s with the re-derivation).
template <int NVW, int T, bool WRITE, bool MASS, bool FROM_Q>
__global__ void __launch_bounds__(SKQ_THREADS, 2) skq_final_kernel(const SkqParams p) {
  extern __shared__ __align__(16) unsigned char skq_smem[];
  constexpr int NG = 2 * NVW;
  const SkqSmem s = skq_smem_setup(skq_smem, p);
  const int warp = threadIdx.x >> 5, lane = lane_id();
  if (warp == SKQ_CW) {
    if (FROM_Q) skq_produce_q32<T>(p, s);
    else skq_produce_dist<T>(p, s);
    return;
  }
  const float bin = *p.bin_score;
  const int rid = skq_rid<T>(lane);
  int c0s[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) c0s[k] = 4 * ((warp * NG + k) * 32 + lane);
  int n = 0, slot = 0;  // running batch counter (matches the producer's) and its ring position
  uint32_t phase = 0;
  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
  SkqCta c;
  if (!skq_item(p, item, c)) continue;
  const int b = c.b;
  float4 v[NG], acc[NG], cbv[NG];
  int4 cbi[NG];
  bool interior[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    interior[k] = c0s[k] + 3 < c.C - 1;
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0s[k] < c.C) {
      v[k] = make_float4(1.f, c0s[k] + 1 < c.C ? 1.f : 0.f, c0s[k] + 2 < c.C ? 1.f : 0.f, c0s[k] + 3 < c.C ? 1.f : 0.f);
      if (p.do_iter) v[k] = v_from_colsum(p.col_prev + (long long)b * p.ldc, c0s[k], c.C);
    }
    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    cbv[k] = make_float4(-1.f, -1.f, -1.f, -1.f);
    cbi[k] = make_int4(0, 0, 0, 0);
  }

  const int nbatch = (c.nrows + T - 1) / T;
  for (int jb = 0; jb < nbatch; ++jb, ++n, skq_ring_next(slot, phase, p.nslots)) {
    const int nb = min(T, c.nrows - jb * T);
    const int i0 = c.row0 + jb * T;
    // per-row scalars: lane t fetches those of row t
    float my_m = 0.f, my_inv = 0.f, my_u = 1.f;
    if (lane < nb) {
      const long long o = (long long)b * p.Rmax + i0 + lane;
      if (!FROM_Q) {
        my_m = p.row_m[o];
        my_inv = p.row_inv[o];
      }
      if (p.do_iter) my_u = p.u[o];
    }
    skq_wait(&s.full_bar[slot], phase);
    const unsigned char* sb = s.ring + (size_t)slot * p.slot_bytes;
    float4 x[T][NG];
#pragma unroll
    for (int t = 0; t < T; ++t)
      if (t < nb) {
        const bool bin_row = (i0 + t == c.R - 1);
        const float* srow = reinterpret_cast<const float*>(sb + (size_t)t * p.row_bytes);
#pragma unroll
        for (int k = 0; k < NG; ++k) {
          if (FROM_Q) x[t][k] = c0s[k] < c.Cq ? *reinterpret_cast<const float4*>(srow + c0s[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
          else x[t][k] = skq_logits_fast(srow, c0s[k], c.C, interior[k], bin_row, bin);
        }
      }

    SkqBest rb[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float mL = skq_max_log2e(__shfl_sync(0xffffffffu, my_m, t));
      const float inv = __shfl_sync(0xffffffffu, my_inv, t);
      const float ui = __shfl_sync(0xffffffffu, my_u, t);
      rb[t].v = -1.f;
      rb[t].j = 0x7fffffff;
      rb[t].m = 0.f;
      if (t < nb) {
        const int i = i0 + t;
        const bool inner_row = i < c.R - 1;
        float* prow = p.P + b * p.p_bs + (long long)i * p.ldp;
#pragma unroll
        for (int k = 0; k < NG; ++k) {
          // (p u) v like the reference's p * u * v, with p = e * (1 / sum); groups beyond C evaluate to exactly 0
          const float pp[4] = {FROM_Q ? x[t][k].x : __fmul_rn(skq_exp(x[t][k].x, mL), inv),
                               FROM_Q ? x[t][k].y : __fmul_rn(skq_exp(x[t][k].y, mL), inv),
                               FROM_Q ? x[t][k].z : __fmul_rn(skq_exp(x[t][k].z, mL), inv),
                               FROM_Q ? x[t][k].w : __fmul_rn(skq_exp(x[t][k].w, mL), inv)};
          const float o[4] = {__fmul_rn(__fmul_rn(pp[0], ui), v[k].x), __fmul_rn(__fmul_rn(pp[1], ui), v[k].y),
                              __fmul_rn(__fmul_rn(pp[2], ui), v[k].z), __fmul_rn(__fmul_rn(pp[3], ui), v[k].w)};
          if (WRITE && c0s[k] < c.C) *reinterpret_cast<float4*>(prow + c0s[k]) = make_float4(o[0], o[1], o[2], o[3]);
          if (inner_row) {
            // column trackers: rows ascend, so a strict > keeps the lowest row (the dustbin column is never flushed)
            if (o[0] > cbv[k].x) { cbv[k].x = o[0]; cbi[k].x = i; }
            if (o[1] > cbv[k].y) { cbv[k].y = o[1]; cbi[k].y = i; }
            if (o[2] > cbv[k].z) { cbv[k].z = o[2]; cbi[k].z = i; }
            if (o[3] > cbv[k].w) { cbv[k].w = o[3]; cbi[k].w = i; }
            if (MASS) {
              acc[k].x += o[0];
              acc[k].y += o[1];
              acc[k].z += o[2];
              acc[k].w += o[3];
            }
            // row tracker over the non-dustbin columns: columns ascend with q and k, strict > keeps the lowest
            if (interior[k]) {  // no column masking
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                rb[t].m += o[q];
                if (o[q] > rb[t].v) {
                  rb[t].v = o[q];
                  rb[t].j = c0s[k] + q;
                }
              }
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const bool in = c0s[k] + q < c.C - 1;
                const float oq = in ? o[q] : -1.f;
                rb[t].m += in ? o[q] : 0.f;
                if (oq > rb[t].v) {
                  rb[t].v = oq;
                  rb[t].j = c0s[k] + q;
                }
              }
            }
          }
        }
      }
    }
    // rows: warp-level transposed reduction, then one (value, column, mass) triple per warp and row through smem
    float* part = s.part + (n & 1) * (3 * T * SKQ_CW);
    {
      const SkqBest r = skq_tr_reduce_best<T>(rb, lane);
      skq_release(&s.empty_bar[slot], lane, r.m);  // every lane's logits sit in registers (they fed the row masses)
      if ((lane & (32 / T - 1)) == 0) {
        part[rid * SKQ_CW + warp] = r.v;
        part[T * SKQ_CW + rid * SKQ_CW + warp] = __int_as_float(r.j);
        part[2 * T * SKQ_CW + rid * SKQ_CW + warp] = r.m;
      }
    }
    skq_consumer_sync();
    if (warp == 0 && lane < nb && i0 + lane < c.R - 1) {
      SkqBest r;
      r.v = -1.f;
      r.j = 0x7fffffff;
      r.m = 0.f;
#pragma unroll
      for (int w = 0; w < SKQ_CW; ++w) {
        SkqBest o;
        o.v = part[lane * SKQ_CW + w];
        o.j = __float_as_int(part[T * SKQ_CW + lane * SKQ_CW + w]);
        o.m = part[2 * T * SKQ_CW + lane * SKQ_CW + w];
        r = skq_best_merge(r, o);
      }
      const long long o = (long long)b * p.N0max + i0 + lane;
      p.row_max[o] = r.v;
      p.row_arg[o] = r.j;
      if (MASS && p.row_mass) p.row_mass[o] = r.m;
    }
  }

  unsigned long long* ck = p.col_key + (long long)b * p.N1max;
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    const float bv[4] = {cbv[k].x, cbv[k].y, cbv[k].z, cbv[k].w};
    const int bi[4] = {cbi[k].x, cbi[k].y, cbi[k].z, cbi[k].w};
    const float am[4] = {acc[k].x, acc[k].y, acc[k].z, acc[k].w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = c0s[k] + q;
      if (j < c.C - 1 && bv[q] >= 0.f) {
        atomicMax(ck + j, pack_max_key(bv[q], bi[q]));
        if (MASS && p.col_mass) atomicAdd(p.col_mass + (long long)b * p.N1max + j, am[q]);
      }
    }
  }
  }  // items
}

// ---------------------------------------------------------------------------------------------------------------
template <int FMT, int NVW, int T_ITER, int T_DIST>
static int run_compact(const SinkhornArgs& a, cudaStream_t st) {
  constexpr int BPE = QFmt<FMT>::BPE;
  constexpr bool FQ = FMT == QF32;                  // final pass streams the exact fp32 copy instead of dist
  constexpr int T_FIN = FQ ? T_ITER : T_DIST;
  const int R = a.N0max + 1, C = a.N1max + 1;
  const int ldq = (C + 15) & ~15;
  const size_t q_row = (size_t)ldq * BPE;
  const size_t d_row = (size_t)((C + 3) & ~3) * 4;  // the padded row (C columns) of the dist passes
  IMP_REQUIRE(a.q_batch_stride % 16 == 0 && (size_t)a.q_batch_stride >= (size_t)R * q_row &&
                  (reinterpret_cast<uintptr_t>(a.q_store) & 15) == 0,
              "sinkhorn: q_store needs %zu bytes per matrix (16-byte aligned), got %lld", (size_t)R * q_row,
              (long long)a.q_batch_stride);
  static DeviceOnce configured;
  if (configured.first()) {
    auto conf = [](const void* f) -> cudaError_t {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, SKQ_SMEM_BUDGET);
      if (e != cudaSuccess) return e;
      return cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    };
    IMP_CUDA_OK(conf((const void*)skq_init_kernel<FMT, NVW, T_DIST>));
    IMP_CUDA_OK(conf((const void*)skq_iter_kernel<FMT, NVW, T_ITER>));
    IMP_CUDA_OK(conf((const void*)skq_final_kernel<NVW, T_FIN, false, false, FQ>));
    IMP_CUDA_OK(conf((const void*)skq_final_kernel<NVW, T_FIN, false, true, FQ>));
    IMP_CUDA_OK(conf((const void*)skq_final_kernel<NVW, T_FIN, true, false, FQ>));
    IMP_CUDA_OK(conf((const void*)skq_final_kernel<NVW, T_FIN, true, true, FQ>));
  }
  SkqParams p;
  p.dist = a.dist; p.dist_bs = a.dist_batch_stride; p.ldd = a.ldd; p.bin_score = a.bin_score;
  p.Q = reinterpret_cast<unsigned char*>(a.q_store); p.q_bs = a.q_batch_stride; p.ldq = ldq; p.Rmax = R;
  p.P = a.P; p.p_bs = a.p_batch_stride; p.ldp = a.ldp;
  p.row_m = a.row_stats; p.row_inv = a.row_stats + (size_t)a.batch * R;
  p.u = a.u; p.ldc = a.ldp;
  p.row_max = a.row_max; p.row_arg = a.row_arg; p.row_mass = a.row_mass; p.col_mass = a.col_mass;
  p.col_key = reinterpret_cast<unsigned long long*>(a.col_key);
  p.n0s = a.n0s; p.n1s = a.n1s; p.N0max = a.N0max; p.N1max = a.N1max;
  p.write_scores = a.write_scores;
  const int iters = a.iters;
  p.do_iter = iters > 0 ? 1 : 0;
  // persistent grid: two CTAs per SM; rows per work item = the multiple of 8 whose item count fills whole rounds of the
  // grid best
  const int ctas = 2 * num_sms();
  int rows_per_cta = sk_rows_per_cta(R, a.batch, ctas, 8);
  if (const char* e = getenv("IMP_SK_ROWS")) {  // tuning override: rows per work item (multiple of 8, 8..128)
    const int r = atoi(e);
    if (r >= 8 && r <= 128 && r % 8 == 0) rows_per_cta = r;
  }
  p.rows_per_cta = rows_per_cta;
  p.blocks_per_mat = (R + rows_per_cta - 1) / rows_per_cta;
  p.n_items = p.blocks_per_mat * a.batch;
  auto slots_for = [&](size_t slot_bytes) {
    int s = (int)((SKQ_SMEM_BUDGET - SKQ_FIXED_SMEM) / slot_bytes);
    return s > 4 ? 4 : s;
  };
  const int slots_d = slots_for(T_DIST * d_row), slots_q = slots_for(T_ITER * q_row);
  IMP_REQUIRE(slots_d >= 2 && slots_q >= 2, "sinkhorn: a row of %d columns does not fit the shared-memory ring", C);
  dim3 grid(p.n_items < ctas ? p.n_items : ctas);
  float* col[3] = {a.colbuf, a.colbuf + (size_t)a.batch * a.ldp, a.colbuf + 2 * (size_t)a.batch * a.ldp};

  p.nslots = slots_d;
  p.row_bytes = (int)d_row;
  p.slot_bytes = (int)(T_DIST * d_row);
  p.col_prev = nullptr;
  p.col_acc = col[0];
  p.col_zero = col[1];
  skq_init_kernel<FMT, NVW, T_DIST><<<grid, SKQ_THREADS, (size_t)slots_d * p.slot_bytes + SKQ_FIXED_SMEM, st>>>(p);

  const bool prof = sk_profiling_on() && iters > 1;
  if (prof) sk_profile_begin(st);
  p.nslots = slots_q;
  p.row_bytes = (int)q_row;
  p.slot_bytes = (int)(T_ITER * q_row);
  for (int k = 1; k < iters; ++k) {
    p.col_prev = col[(k - 1) % 3];
    p.col_acc = col[k % 3];
    p.col_zero = col[(k + 1) % 3];
    skq_iter_kernel<FMT, NVW, T_ITER><<<grid, SKQ_THREADS, (size_t)slots_q * p.slot_bytes + SKQ_FIXED_SMEM, st>>>(p);
  }
  if (prof) sk_profile_end(st, iters - 1);

  if (FQ) {  // slot geometry of the iteration sweeps (rows of the stored copy)
    p.nslots = slots_q;
    p.row_bytes = (int)q_row;
    p.slot_bytes = (int)(T_ITER * q_row);
  } else {
    p.nslots = slots_d;
    p.row_bytes = (int)d_row;
    p.slot_bytes = (int)(T_DIST * d_row);
  }
  p.col_prev = col[(iters > 0 ? iters - 1 : 0) % 3];
  p.col_acc = nullptr;
  p.col_zero = nullptr;
  {
    const size_t smem = (size_t)p.nslots * p.slot_bytes + SKQ_FIXED_SMEM;
    const bool mass = a.col_mass != nullptr || a.row_mass != nullptr;
    if (a.write_scores) {
      if (mass) skq_final_kernel<NVW, T_FIN, true, true, FQ><<<grid, SKQ_THREADS, smem, st>>>(p);
      else skq_final_kernel<NVW, T_FIN, true, false, FQ><<<grid, SKQ_THREADS, smem, st>>>(p);
    } else {
      if (mass) skq_final_kernel<NVW, T_FIN, false, true, FQ><<<grid, SKQ_THREADS, smem, st>>>(p);
      else skq_final_kernel<NVW, T_FIN, false, false, FQ><<<grid, SKQ_THREADS, smem, st>>>(p);
    }
  }
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int FMT>
static int dispatch_nv(const SinkhornArgs& a, cudaStream_t st) {
  const int C = a.N1max + 1;
  if (C <= 2048) return run_compact<FMT, 1, (FMT == QF32 ? 4 : 8), 4>(a, st);
  if (C <= 4096) return run_compact<FMT, 2, (FMT == QF32 ? 2 : 4), 2>(a, st);  // 16 KB rows: fewer rows per ring slot
  set_error("sinkhorn: N1 = %d exceeds the supported maximum of %d columns", a.N1max, 4095);
  return 2;
}

// geometry query (no GPU needed): rows per work item the streaming kernels use for `batch` matrices of N0max + 1 rows on a
// persistent grid of `resident_ctas` CTAs
int sinkhorn_rows_per_item(int batch, int N0max, int resident_ctas) { return sk_rows_per_cta(N0max + 1, batch, resident_ctas, 8); }

int run_sinkhorn_compact(const SinkhornArgs& a, cudaStream_t st) {
  const int storage = a.storage & ~IMP_SK_NO_RESIDENT;
  if (storage == IMP_SK_STORE_F32) return dispatch_nv<QF32>(a, st);
  if (storage == IMP_SK_STORE_F16) return dispatch_nv<QF16>(a, st);
  if (storage == IMP_SK_STORE_F24) return dispatch_nv<QF24>(a, st);
  set_error("sinkhorn: unknown storage format %d", a.storage);
  return 2;
}

}  // namespace imp
