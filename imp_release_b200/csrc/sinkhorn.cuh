#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/imp_b200.h"

namespace imp {

using SinkhornArgs = imp_sinkhorn_args;
using MatchArgs = imp_match_args;

int launch_sinkhorn(const SinkhornArgs& a, cudaStream_t st);
int launch_matches(const MatchArgs& m, cudaStream_t st);

}  // namespace imp
