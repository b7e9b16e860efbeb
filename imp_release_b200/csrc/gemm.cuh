// Split-precision tensor-core GEMM for sm_100a:  D[M,N] = alpha * A[M,K] . B[N,K]^T (+ bias) (+ residual)
//
// Operands live in HBM as fp16 "hi/lo" plane pairs (x = hi + lo carries ~22 mantissa bits).  With
// nsplit = 3 the kernel issues hi.hi + lo.hi + hi.lo per K step (fp32-level accuracy at 1.5x the cost of
// a tf32 GEMM); nsplit = 1 uses the hi planes only.  Reference ops this replaces: every Conv1d(k=1)
// of nets/layers.py (:119 QKV, :134 merge, :149/:218 MLP) and the score einsum of nets/gm.py:290-295.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/imp_b200.h"

namespace imp {

enum GemmOut : int {
  GEMM_OUT_F32 = IMP_GEMM_OUT_F32,                  // fp32 matrix
  GEMM_OUT_F16 = IMP_GEMM_OUT_F16,                  // single fp16 plane
  GEMM_OUT_SPLIT = IMP_GEMM_OUT_SPLIT,              // fp16 hi/lo planes
  GEMM_OUT_SPLIT_RESID = IMP_GEMM_OUT_SPLIT_RESID,  // hi/lo planes of (acc + bias + residual)
};
using GemmArgs = imp_gemm_args;

int launch_gemm(const GemmArgs& g, cudaStream_t stream);

}  // namespace imp
