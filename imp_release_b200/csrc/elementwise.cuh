#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace imp {

int launch_split_planes(const float* x, const float* addend, void* hi, void* lo, long long n, cudaStream_t st);
int launch_merge_planes(const void* hi, const void* lo, float* x, long long n, cudaStream_t st);
int launch_instnorm_relu_split(const float* H, long long h_bs, int ldh, const int* ns, int Nmax, int C, int batch,
                               float eps, int relu, void* out_hi, void* out_lo, float* out_f32, long long o_bs,
                               int ldo, cudaStream_t st);
int launch_instnorm_apply(const float* H, const float* stat_partial, const float* stat_straddle, const int* ns, int Np, int C,
                          int images, float eps, int relu, float* stats, void* out_hi, void* out_lo, cudaStream_t st);
int launch_small_linear(const float* X, int ldx, const float* W, const float* bias, float* Y, int ldy, long long rows,
                        int Cin, int Cout, cudaStream_t st);
int launch_kenc_input(const float* kpts, const float* scores, float* out, long long T, cudaStream_t st);
int launch_gather_rows(const void* in, long long in_bs_bytes, int row_bytes_in, const int* ids, int ids_ld,
                       const int* cnt, void* out, long long out_bs_bytes, int row_bytes_out, int copy_bytes,
                       int max_rows, int batch, cudaStream_t st);

}  // namespace imp
