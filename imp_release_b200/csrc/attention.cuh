#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/imp_b200.h"

namespace imp {

// Multi-head attention over keypoint tokens (MultiHeadedAttention.forward, nets/layers.py:109-136, minus the
// projections): per image `img` (query side) and head h,
//   O = softmax(Q_h K_h^T / 8) V_h        keys restricted to the first nk[src] rows of the source image
// Q/K/V: fp16 [n_img, Nmax, 256], heads contiguous (head h = columns 64h..64h+63; the reference's
// channel-interleaved heads are de-interleaved once in the packed projection weights).
// `shared` mode (SharedAttentionalPropagation, nets/layers.py:211-217): probabilities are re-materialised from the
// stashed Q, K and per-row log-sum-exp of the previous iteration and applied to a new V.
using AttnArgs = imp_attn_args;

int launch_attention(const AttnArgs& a, cudaStream_t st);

// Attention received per source token (EIMP pruning, nets/adgm.py:424-427 / 557-560):
//   colsum[src, m] = sum_h sum_n softmax(Q_h K_h^T / 8)[n, m], recomputed from stashed Q, K and LSE.
using AttnColsumArgs = imp_attn_colsum_args;
int launch_attention_colsum(const AttnColsumArgs& a, cudaStream_t st);

}  // namespace imp
