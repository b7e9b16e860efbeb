// Flash-style multi-head attention for sm_100a (see attention.cuh).  One CTA = 128 query rows of one head of one
// image; keys stream through in tiles of 128.
//   warp 0    : TMA producer (Q once; K/V tiles into a 4-deep ring, 128B swizzle)
//   warp 1    : UMMA issuer  S = Q K^T (SS, fp16, fp32 accum in TMEM) and O += P V (P from TMEM, V consumed
//               token-major as an MN-major B operand); also owns the TMEM allocation
//   warps 2-5 : softmax.  TMEM lane == query row, so a thread owns a whole score row: row max / sum need no
//               shuffles.  exp2 with the 1/8*log2(e) scale folded into one FFMA, lazy rescaling of O (only when the
//               running max grows by > 2^8), P written back to TMEM as packed fp16 over the S columns.
// The kernel is bound by the softmax (MUFU ex2 + issue slots), not by the tensor pipe, so a CTA keeps ONE score
// buffer (256 TMEM columns, 112 KB smem) and TWO CTAs share an SM: while one CTA's softmax warps work, the other
// CTA's MMAs run, and each SM sub-partition always has two softmax warps to switch between.
#include "attention.cuh"

#include <math.h>
#include <stdlib.h>

#include "common.h"
// mbarrier waits with a hardware suspend hint: the plain try_wait loop re-polls every few hundred cycles, and ncu showed 27 %
// of the instructions this kernel issues to be such polls (TMA producer, MMA issuer and softmax warps waiting on each other),
// competing with the softmax warps for issue slots -- the kernel's second limiter after the MUFU.  Measured on B200 (64 pairs,
// N = 2000): 0.786 -> 0.762 ms (667 -> 688 TFLOP/s), sharing layers 0.702 -> 0.685 ms.  No effect on the GEMM or the Sinkhorn
// kernels, which keep the plain loop.
#define IMP_MBAR_SUSPEND_NS 10000
#include "ptx.cuh"

namespace imp {

static constexpr int AT_BM = 128;
static constexpr int CS_BN = 128;  // key rows per CTA in the column-sum kernel
static constexpr int AT_D = 64;
static constexpr int AT_HEADS = 4;
static constexpr int AT_C = AT_D * AT_HEADS;
static constexpr int AT_THREADS = 192;
static constexpr int AT_TILE_BYTES = AT_BM * AT_D * 2;  // 16 KB: a 128-row tile of one head
static constexpr uint32_t AT_COL_S = 0;  // score tile at TMEM column 0 (packed fp16 P aliases its first half), O behind it
static constexpr float AT_SCALE_LOG2 = 0.125f * 1.4426950408889634f;
static constexpr float AT_RESCALE_TAU = 8.0f;

struct AttnKernelParams {
  int n_img, src_offset, Nq_max, Nk_max, shared;
  int lse_only;  // SPLIT kernel only: produce just the row LSE (no V, no P, no O): the pass EIMP's pooling statistics need
  const int *nq, *nk;
  float* lse;
  __half *out_hi, *out_lo;
  long long out_img_stride;
};

// SPLIT = true: "high precision" mode.  Q, K, V arrive as fp16 hi/lo planes and P is split into hi/lo as well; both
// contractions issue the 3-product scheme of the GEMMs (hi.hi + lo.hi + hi.lo), which tracks an fp32 attention to ~1e-5
// (needed when the attention is sharply peaked; costs 3x the tensor work, so the kernel becomes tensor-bound).
template <int BN, int STG, int MINB, bool SPLIT>
__global__ void __launch_bounds__(AT_THREADS, MINB)
attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_ql,
                 const __grid_constant__ CUtensorMap tm_kl, const __grid_constant__ CUtensorMap tm_vl,
                 const AttnKernelParams p) {
  static_assert(!SPLIT || BN == 64, "split mode stores P_hi | P_lo in the 64 score columns");
  constexpr int NP = SPLIT ? 2 : 1;                        // planes per operand
  constexpr int KV_BYTES = BN * AT_D * 2;                  // one K (or V) tile plane
  constexpr uint32_t TMEM_COLS = (BN + AT_D <= 128) ? 128 : 256;  // score tile (BN fp32 columns) + O (64)
  constexpr uint32_t COL_O = BN;
  extern __shared__ __align__(1024) uint8_t smem[];  // 128B-swizzled tiles need 1024-byte alignment
  uint8_t* s_q = smem;                        // Q_hi [, Q_lo]
  uint8_t* s_kv = smem + NP * AT_TILE_BYTES;  // stage s: K_hi, V_hi [, K_lo, V_lo], KV_BYTES each
  constexpr int STAGE_BYTES = 2 * NP * KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_kv + STG * STAGE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + STG;
  uint64_t* s_full = kv_empty + STG;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_done = p_full + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_done + 1);

  if (smem_u32(smem) & 1023u) __trap();
  const int warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * AT_BM;
  const int h = blockIdx.y;
  const int img = blockIdx.z;
  const int src = (img + p.src_offset) % p.n_img;
  const int nq = p.nq ? p.nq[img] : p.Nq_max;
  const int nk = p.nk ? p.nk[src] : p.Nk_max;
  if (q0 >= nq) return;
  const int T = (nk + BN - 1) / BN;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < STG; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one() && T > 0) {
      mbar_arrive_expect_tx(q_full, NP * AT_TILE_BYTES);
      tma_load_3d(s_q, &tm_q, q_full, h * AT_D, q0, img);
      if (SPLIT) tma_load_3d(s_q + AT_TILE_BYTES, &tm_ql, q_full, h * AT_D, q0, img);
      for (int j = 0; j < T; ++j) {
        const int s = j % STG;
        mbar_wait(&kv_empty[s], ((j / STG) & 1) ^ 1);
        uint8_t* st = s_kv + s * STAGE_BYTES;
        const bool no_v = SPLIT && p.lse_only;
        mbar_arrive_expect_tx(&kv_full[s], no_v ? STAGE_BYTES / 2 : STAGE_BYTES);
        tma_load_3d(st, &tm_k, &kv_full[s], h * AT_D, j * BN, src);
        if (!no_v) tma_load_3d(st + KV_BYTES, &tm_v, &kv_full[s], h * AT_D, j * BN, src);
        if (SPLIT) {
          tma_load_3d(st + 2 * KV_BYTES, &tm_kl, &kv_full[s], h * AT_D, j * BN, src);
          if (!no_v) tma_load_3d(st + 3 * KV_BYTES, &tm_vl, &kv_full[s], h * AT_D, j * BN, src);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ UMMA issuer
    if (elect_one() && T > 0) {
      constexpr uint32_t idesc_qk = make_idesc(FMT_F16, AT_BM, BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(FMT_F16, AT_BM, AT_D, 0, 1);  // B = V, MN-major
      const uint32_t q_addr = smem_u32(s_q);
      mbar_wait(q_full, 0);
      for (int j = 0; j < T; ++j) {
        const int st = j % STG;
        mbar_wait(&kv_full[st], (j / STG) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(s_kv + st * STAGE_BYTES);
        const uint32_t v_addr = k_addr + KV_BYTES;
        // S = Q K^T.  The tensor pipe executes MMAs in issue order, so this overwrites the score columns only after
        // O += P(j-1) V(j-1), which read P from the same columns, has drained.
#pragma unroll
        for (int kk = 0; kk < AT_D / 16; ++kk) {
          const uint64_t dq = make_smem_desc_sw128(q_addr + kk * 32, 16, 1024);
          const uint64_t dk = make_smem_desc_sw128(k_addr + kk * 32, 16, 1024);
          umma_f16_ss(tmem_base + AT_COL_S, dq, dk, idesc_qk, kk > 0 ? 1u : 0u);
          if (SPLIT) {
            umma_f16_ss(tmem_base + AT_COL_S, make_smem_desc_sw128(q_addr + AT_TILE_BYTES + kk * 32, 16, 1024), dk, idesc_qk, 1u);
            umma_f16_ss(tmem_base + AT_COL_S, dq, make_smem_desc_sw128(k_addr + 2 * KV_BYTES + kk * 32, 16, 1024), idesc_qk, 1u);
          }
        }
        umma_commit(s_full);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        if (!(SPLIT && p.lse_only))
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {  // 16 keys per MMA: 8 packed fp16x2 columns of P, 2 KB of V
          const uint64_t dv = make_smem_desc_sw128(v_addr + kk * 2048, 1024, 1024);
          umma_f16_ts(tmem_base + COL_O, tmem_base + AT_COL_S + kk * 8, dv, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
          if (SPLIT) {  // P_lo lives in score columns [BN/2, BN)
            umma_f16_ts(tmem_base + COL_O, tmem_base + AT_COL_S + BN / 2 + kk * 8, dv, idesc_pv, 1u);
            umma_f16_ts(tmem_base + COL_O, tmem_base + AT_COL_S + kk * 8,
                        make_smem_desc_sw128(v_addr + 2 * KV_BYTES + kk * 2048, 1024, 1024), idesc_pv, 1u);
          }
        }
        umma_commit(&kv_empty[st]);
        umma_commit(o_done);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane_id();
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int qrow = q0 + row;
    const bool row_ok = qrow < nq;
    float m_run = -INFINITY, l_run = 0.f;
    float lse_in = 0.f;
    if (p.shared && row_ok) lse_in = p.lse[((long long)img * AT_HEADS + h) * p.Nq_max + qrow];

    const uint32_t s_addr = tmem_base + lane_off + AT_COL_S;
    for (int j = 0; j < T; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * BN;
      const int nvalid = nk - kbase;  // keys of this tile that exist (>= 128 except for the ragged last tile)
      float neg_ref;
      const bool ragged = nvalid < BN;  // warp-uniform: only the last tile of a ragged key set needs masking
      if (!p.shared) {
        // pass 1: row max (the score row is re-read from TMEM in pass 2 instead of living in 128 registers)
        float mx = -INFINITY;
#pragma unroll
        for (int cb = 0; cb < BN; cb += 32) {
          uint32_t r[32];
          tmem_ld_x32(s_addr + cb, r);
          tmem_wait_ld();
          if (!ragged) {
            float mx1 = -INFINITY;  // two chains of 3-input maxima (FMNMX3): 16 instructions per 32 columns instead of 32
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              mx = fmaxf(fmaxf(mx, __uint_as_float(r[c])), __uint_as_float(r[c + 1]));
              mx1 = fmaxf(fmaxf(mx1, __uint_as_float(r[c + 2])), __uint_as_float(r[c + 3]));
            }
            mx = fmaxf(mx, mx1);
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) mx = fmaxf(mx, (cb + c < nvalid) ? __uint_as_float(r[c]) : -INFINITY);
          }
        }
        const float m_new = fmaxf(m_run, mx * AT_SCALE_LOG2);
        const bool need = m_new > m_run + AT_RESCALE_TAU;  // first tile: m_run = -inf
        const bool any = __any_sync(0xffffffffu, need);
        if (j > 0) {
          mbar_wait(o_done, (j - 1) & 1);  // O += P(j-1) V(j-1) has landed
          tc_fence_after();
        }
        if (any) {
          const float alpha = need ? fast_exp2(m_run - m_new) : 1.f;
          if (need) {
            l_run *= alpha;
            m_run = m_new;
          }
          if (j > 0 && !(SPLIT && p.lse_only)) {
            const uint32_t o_addr = tmem_base + lane_off + COL_O;
#pragma unroll
            for (int cb = 0; cb < AT_D; cb += 32) {
              uint32_t o[32];
              tmem_ld_x32(o_addr + cb, o);
              tmem_wait_ld();
#pragma unroll
              for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
              tmem_st_x32(o_addr + cb, o);
            }
          }
        }
        neg_ref = -m_run;
      } else {
        if (j > 0) {
          mbar_wait(o_done, (j - 1) & 1);
          tc_fence_after();
        }
        neg_ref = -lse_in;
      }
      float lsum = 0.f;
      if (SPLIT) {
        // high-precision mode: the whole 64-column score row is read first (P_hi | P_lo overwrite all of it), then
        // p is split into an fp16 pair p = hi + lo (~22 bits)
        uint32_t r[BN];
        tmem_ld_x32(s_addr, r);
        tmem_ld_x32(s_addr + 32, r + 32);
        tmem_wait_ld();
        if (p.lse_only) {  // same exponentials, same summation order as below: the LSE comes out bit-identical
#pragma unroll
          for (int c = 0; c < BN; c += 2) {
            float p0 = fast_exp2(fmaf(__uint_as_float(r[c]), AT_SCALE_LOG2, neg_ref));
            float p1 = fast_exp2(fmaf(__uint_as_float(r[c + 1]), AT_SCALE_LOG2, neg_ref));
            if (ragged) {
              p0 = (c < nvalid) ? p0 : 0.f;
              p1 = (c + 1 < nvalid) ? p1 : 0.f;
            }
            lsum += p0 + p1;
          }
        } else {
        uint32_t ph[BN / 2], pl[BN / 2];
#pragma unroll
        for (int c = 0; c < BN; c += 2) {
          float p0 = fast_exp2(fmaf(__uint_as_float(r[c]), AT_SCALE_LOG2, neg_ref));
          float p1 = fast_exp2(fmaf(__uint_as_float(r[c + 1]), AT_SCALE_LOG2, neg_ref));
          if (ragged) {
            p0 = (c < nvalid) ? p0 : 0.f;
            p1 = (c + 1 < nvalid) ? p1 : 0.f;
          }
          lsum += p0 + p1;
          const __half2 hh = __floats2half2_rn(p0, p1);
          const float2 hf = __half22float2(hh);
          const __half2 ll = __floats2half2_rn(p0 - hf.x, p1 - hf.y);
          ph[c >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
          pl[c >> 1] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        tmem_st_x32(s_addr, ph);
        tmem_st_x32(s_addr + BN / 2, pl);
        }
      } else {
      // pass 2: p = exp2(s * c - ref), packed fp16 written over the score columns (chunk cb lands in columns
      // [cb/2, cb/2+16), which this thread has already consumed)
#pragma unroll
      for (int cb = 0; cb < BN; cb += 32) {
        uint32_t r[32];
        tmem_ld_x32(s_addr + cb, r);
        tmem_wait_ld();
        uint32_t pk[16];
        if (!ragged) {
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(r[c]), AT_SCALE_LOG2, neg_ref));
            const float p1 = fast_exp2(fmaf(__uint_as_float(r[c + 1]), AT_SCALE_LOG2, neg_ref));
            lsum += p0 + p1;
            pk[c >> 1] = pack_half2(p0, p1);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float p0 = fast_exp2(fmaf(__uint_as_float(r[c]), AT_SCALE_LOG2, neg_ref));
            float p1 = fast_exp2(fmaf(__uint_as_float(r[c + 1]), AT_SCALE_LOG2, neg_ref));
            p0 = (cb + c < nvalid) ? p0 : 0.f;
            p1 = (cb + c + 1 < nvalid) ? p1 : 0.f;
            lsum += p0 + p1;
            pk[c >> 1] = pack_half2(p0, p1);
          }
        }
        tmem_st_x16(s_addr + (cb >> 1), pk);
      }
      }
      l_run += lsum;
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(p_full);
    }

    // epilogue: O / l -> fp16 hi/lo planes; LSE for the sharing layers / column sums
    __half* oh = p.out_hi + img * p.out_img_stride + (long long)qrow * AT_C + h * AT_D;
    __half* ol = p.out_lo + img * p.out_img_stride + (long long)qrow * AT_C + h * AT_D;
    float v[AT_D];
    if (SPLIT && p.lse_only) {
      if (T > 0) mbar_wait(o_done, (T - 1) & 1);
      if (row_ok) p.lse[((long long)img * AT_HEADS + h) * p.Nq_max + qrow] = m_run + log2f(l_run);
    } else {
    if (T > 0) {
      mbar_wait(o_done, (T - 1) & 1);
      tc_fence_after();
      uint32_t o[AT_D];
      const uint32_t o_addr = tmem_base + lane_off + COL_O;
      tmem_ld_x32(o_addr, o);
      tmem_ld_x32(o_addr + 32, o + 32);
      tmem_wait_ld();
      const float inv = p.shared ? 1.f : 1.f / l_run;
#pragma unroll
      for (int c = 0; c < AT_D; ++c) v[c] = __uint_as_float(o[c]) * inv;
    } else {
#pragma unroll
      for (int c = 0; c < AT_D; ++c) v[c] = 0.f;
    }
    if (row_ok) {
#pragma unroll
      for (int c = 0; c < AT_D; c += 8) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          __half h0, l0, h1, l1;
          split_f16x2(v[c + 2 * t], h0, l0);
          split_f16x2(v[c + 2 * t + 1], h1, l1);
          __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
          hi[t] = *reinterpret_cast<uint32_t*>(&hh);
          lo[t] = *reinterpret_cast<uint32_t*>(&ll);
        }
        *reinterpret_cast<uint4*>(oh + c) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(ol + c) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      if (!p.shared) p.lse[((long long)img * AT_HEADS + h) * p.Nq_max + qrow] = m_run + log2f(l_run);
    }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ---------------------------------------------------------------------------------------------------------------
template <int BN, int STG, int MINB, bool SPLIT>
static int launch_attention_impl(const AttnArgs& a, cudaStream_t st) {
  CUtensorMap tq, tk, tv, tql, tkl, tvl;
  if (make_tmap_f16_3d(&tq, a.q, AT_C, a.Nq_max, a.n_img, a.q_row_stride, a.q_img_stride, AT_D, AT_BM)) return 3;
  if (make_tmap_f16_3d(&tk, a.k, AT_C, a.Nk_max, a.n_img, a.kv_row_stride, a.kv_img_stride, AT_D, BN)) return 3;
  if (make_tmap_f16_3d(&tv, a.v != nullptr ? a.v : a.k, AT_C, a.Nk_max, a.n_img, a.kv_row_stride, a.kv_img_stride, AT_D, BN)) return 3;
  tql = tq;
  tkl = tk;
  tvl = tv;
  if (SPLIT) {
    if (make_tmap_f16_3d(&tql, a.q_lo, AT_C, a.Nq_max, a.n_img, a.q_row_stride, a.q_img_stride, AT_D, AT_BM)) return 3;
    if (make_tmap_f16_3d(&tkl, a.k_lo, AT_C, a.Nk_max, a.n_img, a.kv_row_stride, a.kv_img_stride, AT_D, BN)) return 3;
    if (make_tmap_f16_3d(&tvl, a.v_lo != nullptr ? a.v_lo : a.k_lo, AT_C, a.Nk_max, a.n_img, a.kv_row_stride, a.kv_img_stride, AT_D, BN)) return 3;
  }
  AttnKernelParams p;
  p.n_img = a.n_img;
  p.src_offset = a.src_offset;
  p.Nq_max = a.Nq_max;
  p.Nk_max = a.Nk_max;
  p.shared = a.shared;
  p.lse_only = (SPLIT && a.out_hi == nullptr) ? 1 : 0;
  p.nq = a.nq;
  p.nk = a.nk;
  p.lse = a.lse;
  p.out_hi = reinterpret_cast<__half*>(a.out_hi);
  p.out_lo = reinterpret_cast<__half*>(a.out_lo);
  p.out_img_stride = a.out_img_stride;
  constexpr int NP = SPLIT ? 2 : 1;
  const size_t smem = NP * AT_TILE_BYTES + STG * 2 * NP * (BN * AT_D * 2) + 256;
  auto kern = attention_kernel<BN, STG, MINB, SPLIT>;
  static DeviceOnce configured;
  if (configured.first()) {
    IMP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // several CTAs per SM need the full shared-memory carve-out
    IMP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  dim3 grid((a.Nq_max + AT_BM - 1) / AT_BM, AT_HEADS, a.n_img);
  kern<<<grid, AT_THREADS, smem, st>>>(tq, tk, tv, tql, tkl, tvl, p);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

static int g_attn_variant = -1;
void attention_set_variant(int v) { g_attn_variant = v; }

int launch_attention(const AttnArgs& a, cudaStream_t st) {
  IMP_REQUIRE(a.n_img > 0 && a.Nq_max > 0 && a.Nk_max > 0, "attention: empty problem");
  IMP_REQUIRE(a.q_row_stride >= AT_C && a.kv_row_stride >= AT_C, "attention: row strides must be >= 256");
  if (a.out_hi == nullptr || a.out_lo == nullptr) {  // LSE-only pass (EIMP pooling statistics)
    IMP_REQUIRE(a.q_lo != nullptr && a.k_lo != nullptr && !a.shared && a.lse != nullptr,
                "attention: the LSE-only mode (out_hi = NULL) exists for the split-precision kernel (q_lo, k_lo) and writes lse");
    return launch_attention_impl<64, 2, 2, true>(a, st);
  }
  if (a.q_lo != nullptr || a.k_lo != nullptr || a.v_lo != nullptr) {
    IMP_REQUIRE(a.q_lo != nullptr && a.k_lo != nullptr && a.v_lo != nullptr, "attention: high-precision mode needs q_lo, k_lo and v_lo");
    return launch_attention_impl<64, 2, 2, true>(a, st);  // 96 KB smem, 128 TMEM columns -> 2 CTAs/SM
  }
  // The kernel is bound by the serial QK -> softmax -> PV chain of a CTA, not by a throughput limit, so more (smaller)
  // CTAs per SM win: 64-key tiles need 128 TMEM columns and 48 KB of smem -> up to 4 CTAs/SM; 128-key tiles -> 2.
  if (g_attn_variant < 0) {
    const char* e = getenv("IMP_ATTN_VARIANT");
    g_attn_variant = e ? atoi(e) : 0;
  }
  switch (g_attn_variant) {
    case 1: return launch_attention_impl<128, 3, 2, false>(a, st);
    case 2: return launch_attention_impl<64, 2, 3, false>(a, st);
    default: return launch_attention_impl<64, 2, 4, false>(a, st);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Column sums of the attention map without materialising it: transposed score tiles S^T = K Q^T (keys on the TMEM
// lanes), p = exp2(S^T c - lse[query]) summed along the row, accumulated over the query tiles of one head.
// SPLIT: K and Q arrive as fp16 hi/lo planes and the scores are formed with the 3-product scheme of the GEMMs (fp32-level
// S, consistent with an LSE computed the same way) -- EIMP's pooling compares these sums with their median, so they have
// to track the reference's fp32 attention map, not the fp16 one the layers may use.
// Determinism: a CTA owns (128 keys, one head) and writes its sums with plain stores into a per-head scratch; a second
// tiny kernel adds the four heads in fixed order (no floating-point atomics anywhere).
static constexpr int CS_THREADS = 192;
static constexpr int CS_STAGES = 4;

struct ColsumKernelParams {
  int n_img, src_offset, Nq_max, Nk_max;
  const int *nq, *nk;
  const float* lse;
  float* partial;  // [n_img, 4, Nk_max]
};

template <bool SPLIT>
__global__ void __launch_bounds__(CS_THREADS, 1)
attention_colsum_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_ql, const __grid_constant__ CUtensorMap tm_kl,
                        const ColsumKernelParams p) {
  constexpr int NP = SPLIT ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_k = smem;                        // the CTA's 128 keys (A operand): K_hi [, K_lo]
  uint8_t* s_q = smem + NP * AT_TILE_BYTES;   // ring of query tiles (B operand): stage = Q_hi [, Q_lo]
  constexpr int Q_STAGE = NP * AT_TILE_BYTES;
  float* s_lse = reinterpret_cast<float*>(s_q + CS_STAGES * Q_STAGE);  // [CS_STAGES][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lse + CS_STAGES * CS_BN);
  uint64_t* k_full = bars;
  uint64_t* q_full = bars + 1;
  uint64_t* q_empty = q_full + CS_STAGES;
  uint64_t* s_full = q_empty + CS_STAGES;  // [2]
  uint64_t* s_free = s_full + 2;           // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5;
  const int k0 = blockIdx.x * CS_BN;
  const int h = blockIdx.y;
  const int img = blockIdx.z;  // query image; keys come from src
  const int src = (img + p.src_offset) % p.n_img;
  const int nq = p.nq ? p.nq[img] : p.Nq_max;
  const int nk = p.nk ? p.nk[src] : p.Nk_max;
  float* dst = p.partial + ((long long)img * AT_HEADS + h) * p.Nk_max;
  if (k0 >= nk) {  // no such keys: their sums are 0 (the scratch is not pre-cleared)
    for (int i = threadIdx.x; i < CS_BN && k0 + i < p.Nk_max; i += CS_THREADS) dst[k0 + i] = 0.f;
    return;
  }
  const int T = (nq + AT_BM - 1) / AT_BM;

  if (warp == 0 && elect_one()) {
    mbar_init(k_full, 1);
    for (int s = 0; s < CS_STAGES; ++s) {
      mbar_init(&q_full[s], 1 + 128);  // TMA tx + the softmax threads that staged the LSE slice
      mbar_init(&q_empty[s], 128);  // every softmax thread is done with the LSE slice (and the MMA with Q)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_free[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (elect_one() && T > 0) {
      mbar_arrive_expect_tx(k_full, NP * AT_TILE_BYTES);
      tma_load_3d(s_k, &tm_k, k_full, h * AT_D, k0, src);
      if (SPLIT) tma_load_3d(s_k + AT_TILE_BYTES, &tm_kl, k_full, h * AT_D, k0, src);
      for (int j = 0; j < T; ++j) {
        const int s = j % CS_STAGES;
        mbar_wait(&q_empty[s], ((j / CS_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[s], Q_STAGE);
        tma_load_3d(s_q + s * Q_STAGE, &tm_q, &q_full[s], h * AT_D, j * AT_BM, img);
        if (SPLIT) tma_load_3d(s_q + s * Q_STAGE + AT_TILE_BYTES, &tm_ql, &q_full[s], h * AT_D, j * AT_BM, img);
      }
    }
  } else if (warp == 1) {
    if (elect_one() && T > 0) {
      constexpr uint32_t idesc = make_idesc(FMT_F16, CS_BN, AT_BM, 0, 0);
      const uint32_t k_addr = smem_u32(s_k);
      mbar_wait(k_full, 0);
      for (int j = 0; j < T; ++j) {
        const int s = j % CS_STAGES;
        mbar_wait(&q_full[s], (j / CS_STAGES) & 1);
        mbar_wait(&s_free[j & 1], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(s_q + s * Q_STAGE);
        const uint32_t d = tmem_base + (j & 1) * AT_BM;
#pragma unroll
        for (int kk = 0; kk < AT_D / 16; ++kk) {
          const uint64_t dk = make_smem_desc_sw128(k_addr + kk * 32, 16, 1024);
          const uint64_t dq = make_smem_desc_sw128(q_addr + kk * 32, 16, 1024);
          umma_f16_ss(d, dk, dq, idesc, kk > 0 ? 1u : 0u);
          if (SPLIT) {
            umma_f16_ss(d, make_smem_desc_sw128(k_addr + AT_TILE_BYTES + kk * 32, 16, 1024), dq, idesc, 1u);
            umma_f16_ss(d, dk, make_smem_desc_sw128(q_addr + AT_TILE_BYTES + kk * 32, 16, 1024), idesc, 1u);
          }
        }
        umma_commit(&s_full[j & 1]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int t128 = (warp - 2) * 32 + lane_id();  // 0..127 index among the softmax threads
    const int row = quarter * 32 + lane_id();      // key row (TMEM lane)
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    float acc = 0.f;
    // stage the LSE slices: thread t128 loads lse[query j*128 + t128] for tile j (prefetch distance = ring depth)
    auto stage_lse = [&](int j) {
      const int s = j % CS_STAGES;
      if (j >= CS_STAGES) mbar_wait(&q_empty[s], ((j / CS_STAGES) & 1) ^ 1);
      const int qi = j * AT_BM + t128;
      s_lse[s * CS_BN + t128] = (qi < nq) ? p.lse[((long long)img * AT_HEADS + h) * p.Nq_max + qi] : INFINITY;
      mbar_arrive(&q_full[s]);
    };
    for (int j = 0; j < T && j < CS_STAGES - 1; ++j) stage_lse(j);
    for (int j = 0; j < T; ++j) {
      if (j + CS_STAGES - 1 < T) stage_lse(j + CS_STAGES - 1);
      const int s = j % CS_STAGES;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem_base + lane_off + (j & 1) * AT_BM;
      const float* lse_t = s_lse + s * CS_BN;
#pragma unroll
      for (int cb = 0; cb < AT_BM; cb += 32) {
        uint32_t r[32];
        tmem_ld_x32(s_addr + cb, r);
        tmem_wait_ld();
        // two-level summation: 2000 queries added one by one would leave ~3e-6 of rounding noise on the sums
        float t0 = 0.f, t1 = 0.f;
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          t0 += fast_exp2(fmaf(__uint_as_float(r[c]), AT_SCALE_LOG2, -lse_t[cb + c]));  // lse = +inf -> 0
          t1 += fast_exp2(fmaf(__uint_as_float(r[c + 1]), AT_SCALE_LOG2, -lse_t[cb + c + 1]));
        }
        acc += t0 + t1;
      }
      tc_fence_before();
      mbar_arrive(&s_free[j & 1]);
      // this thread is done with the LSE slice; the MMA that read the Q tile completed before s_full fired
      mbar_arrive(&q_empty[s]);
    }
    if (k0 + row < p.Nk_max) dst[k0 + row] = (k0 + row < nk) ? acc : 0.f;
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// colsum[dst_img][m] = partial[img][0][m] + ... + partial[img][3][m] (fixed order); dst_img = the key image when
// by_key_image, else the query image
__global__ void colsum_reduce_heads_kernel(const float* __restrict__ partial, float* __restrict__ colsum, int n_img,
                                           int Nk_max, int src_offset, int by_key_image) {
  const int img = blockIdx.y;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Nk_max) return;
  const float* pp = partial + (long long)img * AT_HEADS * Nk_max + m;
  const float v = ((pp[0] + pp[Nk_max]) + pp[2 * Nk_max]) + pp[3 * Nk_max];
  const int dst = by_key_image ? (img + src_offset) % n_img : img;
  colsum[(long long)dst * Nk_max + m] = v;
}

template <bool SPLIT>
static int launch_colsum_impl(const AttnColsumArgs& a, cudaStream_t st) {
  CUtensorMap tq, tk, tql, tkl;
  if (make_tmap_f16_3d(&tq, a.q, AT_C, a.Nq_max, a.n_img, a.q_row_stride, a.q_img_stride, AT_D, AT_BM)) return 3;
  if (make_tmap_f16_3d(&tk, a.k, AT_C, a.Nk_max, a.n_img, a.kv_row_stride, a.kv_img_stride, AT_D, CS_BN)) return 3;
  tql = tq;
  tkl = tk;
  if (SPLIT) {
    if (make_tmap_f16_3d(&tql, a.q_lo, AT_C, a.Nq_max, a.n_img, a.q_row_stride, a.q_img_stride, AT_D, AT_BM)) return 3;
    if (make_tmap_f16_3d(&tkl, a.k_lo, AT_C, a.Nk_max, a.n_img, a.kv_row_stride, a.kv_img_stride, AT_D, CS_BN)) return 3;
  }
  ColsumKernelParams p;
  p.n_img = a.n_img;
  p.src_offset = a.src_offset;
  p.Nq_max = a.Nq_max;
  p.Nk_max = a.Nk_max;
  p.nq = a.nq;
  p.nk = a.nk;
  p.lse = a.lse;
  p.partial = a.scratch;
  constexpr int NP = SPLIT ? 2 : 1;
  const size_t smem = NP * AT_TILE_BYTES + CS_STAGES * NP * AT_TILE_BYTES + CS_STAGES * CS_BN * 4 + 1024 + 256;
  static DeviceOnce configured;
  if (configured.first())
    IMP_CUDA_OK(cudaFuncSetAttribute(attention_colsum_kernel<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((a.Nk_max + CS_BN - 1) / CS_BN, AT_HEADS, a.n_img);
  attention_colsum_kernel<SPLIT><<<grid, CS_THREADS, smem, st>>>(tq, tk, tql, tkl, p);
  colsum_reduce_heads_kernel<<<dim3((a.Nk_max + 255) / 256, a.n_img), 256, 0, st>>>(a.scratch, a.colsum, a.n_img, a.Nk_max,
                                                                                    a.src_offset, a.by_key_image);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_attention_colsum(const AttnColsumArgs& a, cudaStream_t st) {
  IMP_REQUIRE(a.n_img > 0 && a.Nq_max > 0 && a.Nk_max > 0, "attention_colsum: empty problem");
  IMP_REQUIRE(a.scratch != nullptr, "attention_colsum: needs a [n_img, 4, Nk_max] float scratch buffer");
  IMP_REQUIRE((a.q_lo == nullptr) == (a.k_lo == nullptr), "attention_colsum: q_lo and k_lo go together");
  return a.q_lo != nullptr ? launch_colsum_impl<true>(a, st) : launch_colsum_impl<false>(a, st);
}

}  // namespace imp
