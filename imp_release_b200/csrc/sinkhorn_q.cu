// Sinkhorn with compact storage of softmax(M) for the iteration sweeps (nets/layers.py:27-46, same recurrence as
// sinkhorn.cu).  At 64 pairs x 2001^2 the 19 iteration sweeps are pure HBM streaming, so the only way to make them
// faster is to move fewer bytes: the init pass writes p = softmax_rows(pad(dist)) as a 16-bit (p * 2^14 in IEEE fp16)
// or 24-bit (top 16 bits of the fp32 word + one byte of mantissa extension, planar) copy, the iteration kernel streams
// that copy through the same shared-memory row ring (1-D bulk TMA copies, producer warp + consumer warps, 2 CTAs/SM),
// and the final pass and the column arg-max re-derive p in fp32 from dist and the saved row statistics (max, 1/sum), so
// the rounding of the copy only enters through the scaling vectors u and v (deviations measured in DESIGN.md).
// All arithmetic on the decoded values is fp32.
//
// Sweep direction alternates from launch to launch (blocks are walked back to front on odd launches) so that the tail
// of the previous sweep, still resident in the 126 MB L2, is what the next sweep reads first.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.h"
#include "sinkhorn_common.cuh"

namespace imp {

int launch_sk_colmax_scaled(const float* P, long long p_bs, int ldp, unsigned long long* col_key, const int* n0s,
                            const int* n1s, int N0max, int N1max, int batch, cudaStream_t st);  // sinkhorn.cu

static constexpr int SKQ_CONSUMERS = 4;
static constexpr int SKQ_THREADS = (SKQ_CONSUMERS + 1) * 32;  // + one producer warp
static constexpr int SKQ_SMEM_BUDGET = 113 * 1024;            // two CTAs per SM
static constexpr float SKQ_F16_SCALE = 16384.f;               // p <= 1 -> <= 2^14; fp16 normals reach down to p = 3.7e-9
static constexpr float SKQ_F16_INV = 1.f / 16384.f;

enum { QF16 = IMP_SK_STORE_F16, QF24 = IMP_SK_STORE_F24 };

struct SkqParams {
  const float* dist;
  long long dist_bs;
  int ldd;
  const float* bin_score;
  unsigned char* Q;  // per matrix: [Rmax][ldq] 16-bit plane, then (24-bit format) [Rmax][ldq] 8-bit plane
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
  const int *n0s, *n1s;
  int N0max, N1max;
  int rows_per_cta, ring_slots, slot_bytes;
  int do_iter, write_scores, reverse;
};

__device__ __forceinline__ void skq_consumer_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(SKQ_CONSUMERS * 32) : "memory");
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
// p = exp(x - max) * (1 / sum): the one expression every pass uses to (re)derive a probability
__device__ __forceinline__ float skq_prob(float x, float m, float inv) { return __fmul_rn(sk_exp(x - m), inv); }

// 24-bit encoding of a non-negative fp32: the top 16 bits verbatim, the low 16 bits L rounded to the nearest multiple of
// 257 (q = (L + 128) / 257 in 0..255, no carry since 255 * 257 = 65535) so that a single byte-permute rebuilds the word
// as [b3 b2 q q]; |error| <= 128.5 ulp(fp32) = 1.5e-5 relative.
__device__ __forceinline__ void skq_enc24(float x, uint32_t& hi, uint32_t& lo) {
  const uint32_t bits = __float_as_uint(x);
  hi = bits >> 16;
  lo = ((bits & 0xFFFFu) + 128u) / 257u;
}

template <int FMT>
__device__ __forceinline__ void skq_store4(unsigned char* qhi, unsigned char* qlo, int c0, float4 t) {
  if (FMT == QF16) {
    uint2 e;
    e.x = pack_half2(t.x * SKQ_F16_SCALE, t.y * SKQ_F16_SCALE);
    e.y = pack_half2(t.z * SKQ_F16_SCALE, t.w * SKQ_F16_SCALE);
    *reinterpret_cast<uint2*>(qhi + 2 * c0) = e;
  } else {
    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
    skq_enc24(t.x, h0, l0);
    skq_enc24(t.y, h1, l1);
    skq_enc24(t.z, h2, l2);
    skq_enc24(t.w, h3, l3);
    *reinterpret_cast<uint2*>(qhi + 2 * c0) = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
    *reinterpret_cast<uint32_t*>(qlo + c0) = l0 | (l1 << 8) | (l2 << 16) | (l3 << 24);
  }
}

// eight consecutive elements of a ring slot -> fp32 (fp16 format: still scaled by 2^14)
template <int FMT>
__device__ __forceinline__ void skq_decode8(const unsigned char* srow, int lo_off, int c0, float (&f)[8]) {
  const uint4 h = *reinterpret_cast<const uint4*>(srow + 2 * c0);
  if (FMT == QF16) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&h.z));
    const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&h.w));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  } else {
    const uint2 l = *reinterpret_cast<const uint2*>(srow + lo_off + c0);
    f[0] = __uint_as_float(__byte_perm(h.x, l.x, 0x1044));
    f[1] = __uint_as_float(__byte_perm(h.x, l.x, 0x3255));
    f[2] = __uint_as_float(__byte_perm(h.y, l.x, 0x1066));
    f[3] = __uint_as_float(__byte_perm(h.y, l.x, 0x3277));
    f[4] = __uint_as_float(__byte_perm(h.z, l.y, 0x1044));
    f[5] = __uint_as_float(__byte_perm(h.z, l.y, 0x3255));
    f[6] = __uint_as_float(__byte_perm(h.w, l.y, 0x1066));
    f[7] = __uint_as_float(__byte_perm(h.w, l.y, 0x3277));
  }
}

// The iteration kernel reads v and accumulates column sums in groups of eight columns per lane; to keep the float4
// shared-memory accesses of a warp on consecutive 16-byte words, columns 8g..8g+3 live at [4g..4g+3] and columns
// 8g+4..8g+7 at [half + 4g ..].
__device__ __forceinline__ int skq_perm(int c, int half) { return ((c >> 3) << 2) + (c & 3) + ((c & 4) ? half : 0); }

struct SkqCta {
  int b, row0, nrows, R, C, Cq;
};
__device__ __forceinline__ bool skq_cta(const SkqParams& p, SkqCta& c) {
  const int bx = p.reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  c.b = p.reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const SkDims d = sk_dims(p.n0s, p.n1s, c.b, p.N0max, p.N1max);
  c.R = d.R;
  c.C = d.C;
  c.Cq = (d.C + 15) & ~15;
  c.row0 = bx * p.rows_per_cta;
  if (c.row0 >= d.R) return false;
  c.nrows = min(p.rows_per_cta, d.R - c.row0);
  return true;
}

// producer warp for the passes that stream dist rows (init, final)
__device__ __forceinline__ void skq_produce_dist(const SkqParams& p, const SkqCta& c, unsigned char* ring, uint64_t* full_bar,
                                                 uint64_t* empty_bar) {
  if (lane_id() != 0) return;
  const float* src = p.dist + c.b * p.dist_bs;
  const uint32_t bytes = (uint32_t)(((c.C - 1 + 3) & ~3) * 4);
  const int S = p.ring_slots;
  for (int r = 0; r < c.nrows; ++r) {
    const int s = r % S;
    mbar_wait(&empty_bar[s], ((r / S) & 1) ^ 1);
    const int i = c.row0 + r;
    if (i == c.R - 1 || bytes == 0) {
      mbar_arrive(&full_bar[s]);  // the dustbin row has no source: the consumer synthesises it
    } else {
      mbar_arrive_expect_tx(&full_bar[s], bytes);
      bulk_copy_g2s(ring + (size_t)s * p.slot_bytes, src + (long long)i * p.ldd, bytes, &full_bar[s]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// init: p = softmax_rows(pad(dist)) -> compact copy + row statistics (+ first half-iteration: u with v = 1 and the
// column sums with that u, from the exact fp32 p)
template <int NV4, int FMT>
__global__ void __launch_bounds__(SKQ_THREADS, 2) skq_init_kernel(const SkqParams p) {
  extern __shared__ __align__(16) unsigned char skq_smem[];
  SkqCta c;
  if (!skq_cta(p, c)) return;
  const int S = p.ring_slots;
  unsigned char* ring = skq_smem;
  float* s_col = reinterpret_cast<float*>(ring + (size_t)S * p.slot_bytes) + p.ldq;  // [ldq] (after the unused s_v)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_col + p.ldq);
  uint64_t* empty_bar = full_bar + S;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == SKQ_CONSUMERS) {
    skq_produce_dist(p, c, ring, full_bar, empty_bar);
    return;
  }
  const int ct = threadIdx.x;
  const int b = c.b;
  if (c.row0 == 0 && p.col_zero != nullptr)
    for (int j = ct; j < p.ldc; j += SKQ_CONSUMERS * 32) p.col_zero[(long long)b * p.ldc + j] = 0.f;
  for (int j = ct; j < p.ldq; j += SKQ_CONSUMERS * 32) s_col[j] = 0.f;
  skq_consumer_sync();

  const float bin = *p.bin_score;
  float4 acc[NV4];
#pragma unroll
  for (int k = 0; k < NV4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int r = warp; r < c.nrows; r += SKQ_CONSUMERS) {
    const int s = r % S;
    const int i = c.row0 + r;
    mbar_wait(&full_bar[s], (r / S) & 1);
    float* srow = reinterpret_cast<float*>(ring + (size_t)s * p.slot_bytes);
    const bool bin_row = (i == c.R - 1);
    // pass 1: materialise the padded logits in the slot and take the row max (each lane only touches its own groups)
    float m = -FLT_MAX;
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
      const int c0 = 4 * (lane_id() + 32 * k);
      if (c0 < c.C) {
        const float4 t = skq_logits(srow, c0, c.C, bin_row, bin);
        if (bin_row || c0 + 3 >= c.C - 1) *reinterpret_cast<float4*>(srow + c0) = t;
        m = fmaxf(m, fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)));
      }
    }
    m = warp_max(m);
    // pass 2: e = exp(x - max) in place, row sum
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
      const int c0 = 4 * (lane_id() + 32 * k);
      if (c0 < c.C) {
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
    const float ui = p.do_iter ? (bin_row ? (float)c.R : 1.f) / (sum * inv_sum + SK_EPS) : 0.f;
    if (lane_id() == 0) {
      if (p.do_iter) p.u[(long long)b * p.Rmax + i] = ui;
      p.row_m[(long long)b * p.Rmax + i] = m;
      p.row_inv[(long long)b * p.Rmax + i] = inv_sum;
    }
    // pass 3: normalise, encode, store; fold p * u into the column accumulators
    unsigned char* qhi = p.Q + b * p.q_bs + (size_t)i * p.ldq * 2;
    unsigned char* qlo = p.Q + b * p.q_bs + (size_t)p.Rmax * p.ldq * 2 + (size_t)i * p.ldq;
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
      const int c0 = 4 * (lane_id() + 32 * k);
      if (c0 < c.C) {
        float4 t = *reinterpret_cast<const float4*>(srow + c0);
        t.x = __fmul_rn(t.x, inv_sum);
        t.y = __fmul_rn(t.y, inv_sum);
        t.z = __fmul_rn(t.z, inv_sum);
        t.w = __fmul_rn(t.w, inv_sum);
        skq_store4<FMT>(qhi, qlo, c0, t);
        acc[k].x += t.x * ui;
        acc[k].y += t.y * ui;
        acc[k].z += t.z * ui;
        acc[k].w += t.w * ui;
      } else if (c0 < c.Cq) {
        skq_store4<FMT>(qhi, qlo, c0, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
    __syncwarp();
    if (lane_id() == 0) mbar_arrive(&empty_bar[s]);
  }

  if (p.do_iter) {
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
      const int c0 = 4 * (lane_id() + 32 * k);
      if (c0 < c.C) {
        atomicAdd(s_col + c0 + 0, acc[k].x);
        atomicAdd(s_col + c0 + 1, acc[k].y);
        atomicAdd(s_col + c0 + 2, acc[k].z);
        atomicAdd(s_col + c0 + 3, acc[k].w);
      }
    }
    skq_consumer_sync();
    for (int j = ct; j < c.C; j += SKQ_CONSUMERS * 32) atomicAdd(p.col_acc + (long long)b * p.ldc + j, s_col[j]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// one Sinkhorn iteration in a single sweep over the compact copy: u_i = r_i / (sum_j p_ij v_j + eps), then p_ij u_i
// folded into the column sums the next launch turns into v
template <int NV8, int FMT>
__global__ void __launch_bounds__(SKQ_THREADS, 2) skq_iter_kernel(const SkqParams p) {
  extern __shared__ __align__(16) unsigned char skq_smem[];
  SkqCta c;
  if (!skq_cta(p, c)) return;
  const int S = p.ring_slots;
  unsigned char* ring = skq_smem;
  float* s_v = reinterpret_cast<float*>(ring + (size_t)S * p.slot_bytes);  // [ldq], permuted (skq_perm), pre-scaled
  float* s_col = s_v + p.ldq;                                               // [ldq], permuted
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_col + p.ldq);
  uint64_t* empty_bar = full_bar + S;
  const int warp = threadIdx.x >> 5;
  const int half = p.ldq >> 1;
  const int lo_off = p.ldq * 2;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int b = c.b;

  if (warp == SKQ_CONSUMERS) {
    if (lane_id() == 0) {
      const unsigned char* qhi = p.Q + b * p.q_bs;
      const unsigned char* qlo = qhi + (size_t)p.Rmax * p.ldq * 2;
      const uint32_t hi_bytes = (uint32_t)c.Cq * 2, lo_bytes = (uint32_t)c.Cq;
      for (int r = 0; r < c.nrows; ++r) {
        const int s = r % S;
        mbar_wait(&empty_bar[s], ((r / S) & 1) ^ 1);
        const int i = c.row0 + r;
        unsigned char* dst = ring + (size_t)s * p.slot_bytes;
        mbar_arrive_expect_tx(&full_bar[s], FMT == QF16 ? hi_bytes : hi_bytes + lo_bytes);
        bulk_copy_g2s(dst, qhi + (size_t)i * p.ldq * 2, hi_bytes, &full_bar[s]);
        if (FMT == QF24) bulk_copy_g2s(dst + lo_off, qlo + (size_t)i * p.ldq, lo_bytes, &full_bar[s]);
      }
    }
    return;
  }

  const int ct = threadIdx.x;
  if (c.row0 == 0 && p.col_zero != nullptr)
    for (int j = ct; j < p.ldc; j += SKQ_CONSUMERS * 32) p.col_zero[(long long)b * p.ldc + j] = 0.f;
  const float vscale = (FMT == QF16) ? SKQ_F16_INV : 1.f;
  for (int c0 = 4 * ct; c0 < p.ldq; c0 += 4 * SKQ_CONSUMERS * 32) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 < c.C) v = v_from_colsum(p.col_prev + (long long)b * p.ldc, c0, c.C);  // C <= ldc, both multiples of 4 apart
    v.x *= vscale;
    v.y *= vscale;
    v.z *= vscale;
    v.w *= vscale;
    const int o = skq_perm(c0, half);  // c0 % 4 == 0: the four columns stay adjacent
    *reinterpret_cast<float4*>(s_v + o) = v;
    *reinterpret_cast<float4*>(s_col + o) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  skq_consumer_sync();

  float acc[NV8][8];
#pragma unroll
  for (int k = 0; k < NV8; ++k)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[k][q] = 0.f;

  for (int r = warp; r < c.nrows; r += SKQ_CONSUMERS) {
    const int s = r % S;
    const int i = c.row0 + r;
    mbar_wait(&full_bar[s], (r / S) & 1);
    const unsigned char* srow = ring + (size_t)s * p.slot_bytes;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int k = 0; k < NV8; ++k) {
      const int g = lane_id() + 32 * k;
      if (8 * g < c.Cq) {
        float f[8];
        skq_decode8<FMT>(srow, lo_off, 8 * g, f);
        const float4 va = *reinterpret_cast<const float4*>(s_v + 4 * g);
        const float4 vb = *reinterpret_cast<const float4*>(s_v + half + 4 * g);
        rs0 += (f[0] * va.x + f[1] * va.y) + (f[2] * va.z + f[3] * va.w);
        rs1 += (f[4] * vb.x + f[5] * vb.y) + (f[6] * vb.z + f[7] * vb.w);
      }
    }
    const float rs = warp_sum(rs0 + rs1);
    const float ui = ((i == c.R - 1) ? (float)c.R : 1.f) / (rs + SK_EPS);
    if (lane_id() == 0) p.u[(long long)b * p.Rmax + i] = ui;
    const float uis = ui * vscale;
#pragma unroll
    for (int k = 0; k < NV8; ++k) {
      const int g = lane_id() + 32 * k;
      if (8 * g < c.Cq) {
        float f[8];
        skq_decode8<FMT>(srow, lo_off, 8 * g, f);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[k][q] = fmaf(f[q], uis, acc[k][q]);
      }
    }
    __syncwarp();
    if (lane_id() == 0) mbar_arrive(&empty_bar[s]);  // the ring slot is free again
  }

  // combine the CTA's warps in shared memory, then one global atomic per column
#pragma unroll
  for (int k = 0; k < NV8; ++k) {
    const int g = lane_id() + 32 * k;
    if (8 * g < c.Cq) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        atomicAdd(s_col + 4 * g + q, acc[k][q]);
        atomicAdd(s_col + half + 4 * g + q, acc[k][4 + q]);
      }
    }
  }
  skq_consumer_sync();
  for (int j = ct; j < c.C; j += SKQ_CONSUMERS * 32) atomicAdd(p.col_acc + (long long)b * p.ldc + j, s_col[skq_perm(j, half)]);
}

// ---------------------------------------------------------------------------------------------------------------
// final: out = (p u) v with p re-derived in fp32 from dist and the row statistics; row arg-max / masses over the
// non-dustbin block; scores written to P only when the caller wants the matrix
template <int NV4>
__global__ void __launch_bounds__(SKQ_THREADS, 2) skq_final_kernel(const SkqParams p) {
  extern __shared__ __align__(16) unsigned char skq_smem[];
  SkqCta c;
  if (!skq_cta(p, c)) return;
  const int S = p.ring_slots;
  unsigned char* ring = skq_smem;
  float* s_v = reinterpret_cast<float*>(ring + (size_t)S * p.slot_bytes);  // [ldq], plain layout, unscaled
  float* s_col = s_v + p.ldq;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_col + p.ldq);
  uint64_t* empty_bar = full_bar + S;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == SKQ_CONSUMERS) {
    skq_produce_dist(p, c, ring, full_bar, empty_bar);
    return;
  }
  const int ct = threadIdx.x;
  const int b = c.b;
  const int C4 = (c.C + 3) & ~3;
  for (int c0 = 4 * ct; c0 < C4; c0 += 4 * SKQ_CONSUMERS * 32) {
    float4 v = make_float4(c0 + 0 < c.C ? 1.f : 0.f, c0 + 1 < c.C ? 1.f : 0.f, c0 + 2 < c.C ? 1.f : 0.f, c0 + 3 < c.C ? 1.f : 0.f);
    if (p.do_iter) v = v_from_colsum(p.col_prev + (long long)b * p.ldc, c0, c.C);
    *reinterpret_cast<float4*>(s_v + c0) = v;
    *reinterpret_cast<float4*>(s_col + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  skq_consumer_sync();

  const float bin = *p.bin_score;
  const bool want_col = p.col_mass != nullptr;
  float4 acc[NV4];
#pragma unroll
  for (int k = 0; k < NV4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int r = warp; r < c.nrows; r += SKQ_CONSUMERS) {
    const int s = r % S;
    const int i = c.row0 + r;
    const float ui = p.do_iter ? p.u[(long long)b * p.Rmax + i] : 1.f;
    const float m = p.row_m[(long long)b * p.Rmax + i];
    const float inv = p.row_inv[(long long)b * p.Rmax + i];
    mbar_wait(&full_bar[s], (r / S) & 1);
    const float* srow = reinterpret_cast<const float*>(ring + (size_t)s * p.slot_bytes);
    const bool bin_row = (i == c.R - 1);
    const bool inner_row = !bin_row;
    float* prow = p.P + b * p.p_bs + (long long)i * p.ldp;
    if (inner_row || p.write_scores) {
      // four independent (value, column) trackers -- one per float4 component -- keep the compare/select chains short
      float bv[4] = {-1.f, -1.f, -1.f, -1.f}, ms[4] = {0.f, 0.f, 0.f, 0.f};
      int bj[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
#pragma unroll
      for (int k = 0; k < NV4; ++k) {
        const int c0 = 4 * (lane_id() + 32 * k);
        if (c0 >= c.C) continue;
        const float4 x = skq_logits(srow, c0, c.C, bin_row, bin);
        const float4 v = *reinterpret_cast<const float4*>(s_v + c0);
        const float o[4] = {__fmul_rn(__fmul_rn(skq_prob(x.x, m, inv), ui), v.x), __fmul_rn(__fmul_rn(skq_prob(x.y, m, inv), ui), v.y),
                            __fmul_rn(__fmul_rn(skq_prob(x.z, m, inv), ui), v.z), __fmul_rn(__fmul_rn(skq_prob(x.w, m, inv), ui), v.w)};
        if (p.write_scores) *reinterpret_cast<float4*>(prow + c0) = make_float4(o[0], o[1], o[2], o[3]);
        if (inner_row) {
          if (c0 + 3 < c.C - 1) {  // interior group: no column masking needed
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
              const bool in = c0 + q < c.C - 1;
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
      if (inner_row) {
        float best = bv[0], mass = (ms[0] + ms[1]) + (ms[2] + ms[3]);
        int best_j = bj[0];
#pragma unroll
        for (int q = 1; q < 4; ++q)
          if (bv[q] > best || (bv[q] == best && bj[q] < best_j)) {
            best = bv[q];
            best_j = bj[q];
          }
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
    if (lane_id() == 0) mbar_arrive(&empty_bar[s]);
  }

  if (want_col) {
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
      const int c0 = 4 * (lane_id() + 32 * k);
      if (c0 < c.C) {
        atomicAdd(s_col + c0 + 0, acc[k].x);
        atomicAdd(s_col + c0 + 1, acc[k].y);
        atomicAdd(s_col + c0 + 2, acc[k].z);
        atomicAdd(s_col + c0 + 3, acc[k].w);
      }
    }
    skq_consumer_sync();
    for (int j = ct; j < c.C - 1; j += SKQ_CONSUMERS * 32) atomicAdd(p.col_mass + (long long)b * p.N1max + j, s_col[j]);
  }
}

// column arg-max over the non-dustbin block, scores re-derived from dist exactly like the final pass does:
// thread per column (coalesced), row slabs, packed atomicMax (lowest row wins ties)
__global__ void __launch_bounds__(128)
skq_colmax_kernel(const float* __restrict__ dist, long long dist_bs, int ldd, const float* __restrict__ row_m,
                  const float* __restrict__ row_inv, const float* __restrict__ u, const float* __restrict__ col_last, int ldc,
                  int has_iter, unsigned long long* __restrict__ col_key, const int* __restrict__ n0s,
                  const int* __restrict__ n1s, int N0max, int N1max, int slab, int reverse) {
  const int b = reverse ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
  const int by = reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const SkDims d = sk_dims(n0s, n1s, b, N0max, N1max);
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i0 = by * slab;
  if (j >= d.C - 1 || i0 >= d.R - 1) return;
  const int i1 = min(i0 + slab, d.R - 1);
  const float vj = has_iter ? 1.f / (col_last[(long long)b * ldc + j] + SK_EPS) : 1.f;  // c_j = 1 for j < C-1
  const float* base = dist + b * dist_bs + j;
  const long long rb = (long long)b * (N0max + 1);
  float best = -1.f;
  int bi = 0;
#pragma unroll 4
  for (int i = i0; i < i1; ++i) {
    const float pr = skq_prob(base[(long long)i * ldd], row_m[rb + i], row_inv[rb + i]);
    const float val = __fmul_rn(__fmul_rn(pr, has_iter ? u[rb + i] : 1.f), vj);
    if (val > best) {
      best = val;
      bi = i;
    }
  }
  atomicMax(col_key + (long long)b * N1max + j, pack_max_key(best, bi));
}

template <int NV8, int FMT>
static int run_compact(const SinkhornArgs& a, cudaStream_t st) {
  constexpr int NV4 = 2 * NV8;
  const int R = a.N0max + 1, C = a.N1max + 1;
  const int ldq = (C + 15) & ~15;
  const size_t q_row = (size_t)ldq * (FMT == QF16 ? 2 : 3);
  const size_t d_row = (size_t)((C + 3) & ~3) * 4;  // the padded row (C columns) is materialised in the slot
  const size_t fixed = 2 * (size_t)ldq * sizeof(float) + 2 * 64 * sizeof(uint64_t);
  IMP_REQUIRE(a.q_batch_stride % 16 == 0 && (size_t)a.q_batch_stride >= (size_t)R * q_row &&
                  (reinterpret_cast<uintptr_t>(a.q_store) & 15) == 0,
              "sinkhorn: q_store needs %zu bytes per matrix (16-byte aligned), got %lld", (size_t)R * q_row,
              (long long)a.q_batch_stride);
  static bool configured = false;
  if (!configured) {
    auto conf = [](const void* f) -> cudaError_t {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, SKQ_SMEM_BUDGET);
      if (e != cudaSuccess) return e;
      return cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    };
    IMP_CUDA_OK(conf((const void*)skq_init_kernel<NV4, FMT>));
    IMP_CUDA_OK(conf((const void*)skq_iter_kernel<NV8, FMT>));
    IMP_CUDA_OK(conf((const void*)skq_final_kernel<NV4>));
    configured = true;
  }
  SkqParams p;
  p.dist = a.dist; p.dist_bs = a.dist_batch_stride; p.ldd = a.ldd; p.bin_score = a.bin_score;
  p.Q = reinterpret_cast<unsigned char*>(a.q_store); p.q_bs = a.q_batch_stride; p.ldq = ldq; p.Rmax = R;
  p.P = a.P; p.p_bs = a.p_batch_stride; p.ldp = a.ldp;
  p.row_m = a.row_stats; p.row_inv = a.row_stats + (size_t)a.batch * R;
  p.u = a.u; p.ldc = a.ldp;
  p.row_max = a.row_max; p.row_arg = a.row_arg; p.row_mass = a.row_mass; p.col_mass = a.col_mass;
  p.n0s = a.n0s; p.n1s = a.n1s; p.N0max = a.N0max; p.N1max = a.N1max;
  p.write_scores = a.write_scores;
  const int iters = a.iters;
  p.do_iter = iters > 0 ? 1 : 0;
  const int rows_per_cta = sk_rows_per_cta(R, a.batch, num_sms(), SKQ_CONSUMERS);
  p.rows_per_cta = rows_per_cta;
  auto slots_for = [&](size_t row_bytes) {
    int s = (int)((SKQ_SMEM_BUDGET - fixed) / row_bytes);
    if (s > 64) s = 64;
    if (s > rows_per_cta) s = rows_per_cta;
    // slot s must always be drained by the same consumer warp (row r -> warp r % CONSUMERS, slot r % slots)
    return s / SKQ_CONSUMERS * SKQ_CONSUMERS;
  };
  const int slots_d = slots_for(d_row), slots_q = slots_for(q_row);
  IMP_REQUIRE(slots_d >= SKQ_CONSUMERS && slots_q >= SKQ_CONSUMERS,
              "sinkhorn: a row of %d columns does not fit the shared-memory ring", C);
  dim3 grid((R + rows_per_cta - 1) / rows_per_cta, a.batch);
  float* col[3] = {a.colbuf, a.colbuf + (size_t)a.batch * a.ldp, a.colbuf + 2 * (size_t)a.batch * a.ldp};

  p.ring_slots = slots_d;
  p.slot_bytes = (int)d_row;
  p.col_prev = nullptr;
  p.col_acc = col[0];
  p.col_zero = col[1];
  p.reverse = 0;
  skq_init_kernel<NV4, FMT><<<grid, SKQ_THREADS, slots_d * d_row + fixed, st>>>(p);

  const bool prof = sk_profiling_on() && iters > 1;
  if (prof) sk_profile_begin(st);
  p.ring_slots = slots_q;
  p.slot_bytes = (int)q_row;
  for (int k = 1; k < iters; ++k) {
    p.col_prev = col[(k - 1) % 3];
    p.col_acc = col[k % 3];
    p.col_zero = col[(k + 1) % 3];
    p.reverse = k & 1;
    skq_iter_kernel<NV8, FMT><<<grid, SKQ_THREADS, slots_q * q_row + fixed, st>>>(p);
  }
  if (prof) sk_profile_end(st, iters - 1);

  const float* col_last = col[(iters > 0 ? iters - 1 : 0) % 3];
  p.ring_slots = slots_d;
  p.slot_bytes = (int)d_row;
  p.col_prev = col_last;
  p.col_acc = nullptr;
  p.col_zero = nullptr;
  p.reverse = 0;
  skq_final_kernel<NV4><<<grid, SKQ_THREADS, slots_d * d_row + fixed, st>>>(p);
  if (a.write_scores) {
    if (int rc = launch_sk_colmax_scaled(a.P, a.p_batch_stride, a.ldp, reinterpret_cast<unsigned long long*>(a.col_key), a.n0s,
                                         a.n1s, a.N0max, a.N1max, a.batch, st))
      return rc;
  } else {
    const int slab = 256;
    skq_colmax_kernel<<<dim3((a.N1max + 127) / 128, (a.N0max + slab - 1) / slab, a.batch), 128, 0, st>>>(
        a.dist, a.dist_batch_stride, a.ldd, p.row_m, p.row_inv, a.u, col_last, a.ldp, iters > 0 ? 1 : 0,
        reinterpret_cast<unsigned long long*>(a.col_key), a.n0s, a.n1s, a.N0max, a.N1max, slab, 1);
  }
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int FMT>
static int dispatch_nv(const SinkhornArgs& a, cudaStream_t st) {
  const int C = a.N1max + 1;
  if (C <= 256 * 2) return run_compact<2, FMT>(a, st);
  if (C <= 256 * 4) return run_compact<4, FMT>(a, st);
  if (C <= 256 * 8) return run_compact<8, FMT>(a, st);
  if (C <= 256 * 13) return run_compact<13, FMT>(a, st);
  set_error("sinkhorn: N1 = %d exceeds the supported maximum of %d columns", a.N1max, 256 * 13 - 1);
  return 2;
}

int run_sinkhorn_compact(const SinkhornArgs& a, cudaStream_t st) {
  if (a.storage == IMP_SK_STORE_F16) return dispatch_nv<QF16>(a, st);
  if (a.storage == IMP_SK_STORE_F24) return dispatch_nv<QF24>(a, st);
  set_error("sinkhorn: unknown storage format %d", a.storage);
  return 2;
}

}  // namespace imp
