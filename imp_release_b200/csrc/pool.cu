// EIMP adaptive pooling keep-rule on device (AdaGMN.pool nets/adgm.py:552-605, batched rule nets/adgm.py:463-500):
// threshold on the Sinkhorn row mass, LOWER medians (torch.median) of the normalised received attention over the
// rows that passed, union with the rows above either median, sorted compaction.  One CTA per sample; the medians
// come from a bitonic sort in shared memory (N <= 4096), the compaction from a block prefix sum, so no host sync
// is needed between iterations.
#include "../../include/imp_b200.h"
#include "common.h"
#include "ptx.cuh"

#include <float.h>

namespace imp {

static constexpr int POOL_THREADS = 1024;
static constexpr int POOL_MAXN = 4096;

__device__ float block_sum(float v, float* s_red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if (lane_id() == 0) s_red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < POOL_THREADS / 32) ? s_red[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (lane_id() == 0) s_red[0] = t;
  }
  __syncthreads();
  return s_red[0];
}

// ascending bitonic sort of s[0..n2) (n2 power of two)
__device__ void bitonic_sort(float* s, int n2) {
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += POOL_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float a = s[i], b = s[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            s[i] = b;
            s[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(POOL_THREADS)
pool_select_kernel(const imp_pool_args a) {
  __shared__ float s_val[POOL_MAXN];
  __shared__ float s_red[POOL_THREADS / 32];
  __shared__ int s_scan[POOL_THREADS / 32];
  __shared__ int s_cnt;
  const int b = blockIdx.x;
  const int cnt = a.cnt_in[b];
  const int* ids_in = a.ids_in + (long long)b * a.ld;
  int* ids_out = a.ids_out + (long long)b * a.ld;
  const float* mass = a.mass + (long long)b * a.ld;
  const float* as = a.a_self + (long long)b * a.ld;
  const float* ac = a.a_cross + (long long)b * a.ld;

  // pids and their count
  int local = 0;
  for (int i = threadIdx.x; i < cnt; i += POOL_THREADS) local += (mass[i] >= a.thresh) ? 1 : 0;
  const int npid = (int)(block_sum((float)local, s_red) + 0.5f);
  const bool update = !(a.n_min_tokens > 0 && cnt <= a.n_min_tokens) && npid > 0;
  if (!update) {
    for (int i = threadIdx.x; i < cnt; i += POOL_THREADS) ids_out[i] = ids_in[i];
    if (threadIdx.x == 0) {
      a.cnt_out[b] = cnt;
      a.changed[b] = 0;
    }
    return;
  }
  // normalisation sums (norm_prob = sum_prob / sum(sum_prob), nets/adgm.py:429-432; pruned tokens contribute 0)
  float ls = 0.f, lc = 0.f;
  for (int i = threadIdx.x; i < cnt; i += POOL_THREADS) {
    ls += as[i];
    lc += ac[i];
  }
  const float tot_s = block_sum(ls, s_red);
  const float tot_c = block_sum(lc, s_red);

  int n2 = 1;
  while (n2 < cnt) n2 <<= 1;
  const int kth = (npid - 1) >> 1;  // torch.median = lower median
  float med[2];
  for (int which = 0; which < 2; ++which) {
    const float* arr = which == 0 ? as : ac;
    const float tot = which == 0 ? tot_s : tot_c;
    for (int i = threadIdx.x; i < n2; i += POOL_THREADS)
      s_val[i] = (i < cnt && mass[i] >= a.thresh) ? arr[i] / tot : FLT_MAX;
    __syncthreads();
    bitonic_sort(s_val, n2);
    med[which] = s_val[kth];
    __syncthreads();
  }
  // keep flags + ordered compaction (ids_in is sorted, so the output stays sorted like torch.unique)
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  for (int base = 0; base < cnt; base += POOL_THREADS) {
    const int i = base + threadIdx.x;
    int keep = 0, g = 0;
    if (i < cnt) {
      g = ids_in[i];
      keep = (mass[i] >= a.thresh) || (as[i] / tot_s >= med[0]) || (ac[i] / tot_c >= med[1]);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int w = threadIdx.x >> 5;
    if (lane_id() == 0) s_scan[w] = __popc(bal);
    __syncthreads();
    int off = s_cnt;
    for (int k = 0; k < w; ++k) off += s_scan[k];
    if (keep) ids_out[off + __popc(bal & ((1u << lane_id()) - 1))] = g;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int k = 0; k < POOL_THREADS / 32; ++k) t += s_scan[k];
      s_cnt += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    a.cnt_out[b] = s_cnt;
    a.changed[b] = 1;
  }
}

int launch_pool_select(const imp_pool_args& a, cudaStream_t st) {
  IMP_REQUIRE(a.ld <= POOL_MAXN, "pool_select: at most %d tokens per image supported (got %d)", POOL_MAXN, a.ld);
  if (a.batch == 0) return 0;
  pool_select_kernel<<<a.batch, POOL_THREADS, 0, st>>>(a);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace imp

namespace imp {

__global__ void scatter_matches_kernel(const long long* __restrict__ idx0, const float* __restrict__ ms0, int ld_sub,
                                       const int* __restrict__ gids0, const int* __restrict__ gids1, int ld_ids,
                                       const int* __restrict__ cnt0, long long* __restrict__ out_idx,
                                       float* __restrict__ out_ms, int ld_out) {
  const int b = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= cnt0[b]) return;
  const int g0 = gids0[(long long)b * ld_ids + r];
  const long long j = idx0[(long long)b * ld_sub + r];
  out_ms[(long long)b * ld_out + g0] = ms0[(long long)b * ld_sub + r];
  if (j >= 0) out_idx[(long long)b * ld_out + g0] = gids1[(long long)b * ld_ids + j];
}

int launch_scatter_matches(const long long* idx0, const float* ms0, int ld_sub, const int* gids0, const int* gids1,
                           int ld_ids, const int* cnt0, long long* out_idx, float* out_ms, int ld_out, int batch,
                           cudaStream_t st) {
  if (batch == 0 || ld_sub == 0) return 0;
  scatter_matches_kernel<<<dim3((ld_sub + 255) / 256, batch), 256, 0, st>>>(idx0, ms0, ld_sub, gids0, gids1, ld_ids, cnt0,
                                                                            out_idx, out_ms, ld_out);
  IMP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace imp
