// extern "C" surface of libimp_b200.so (declared in include/imp_b200.h).
#include "../../include/imp_b200.h"

#include "attention.cuh"
#include "common.h"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "sinkhorn.cuh"

namespace imp {
int launch_pool_select(const imp_pool_args& a, cudaStream_t st);
int launch_scatter_matches(const long long* idx0, const float* ms0, int ld_sub, const int* gids0, const int* gids1,
                           int ld_ids, const int* cnt0, long long* out_idx, float* out_ms, int ld_out, int batch,
                           cudaStream_t st);
int launch_score_argmax(const float* P, long long p_bs, int ldp, float* row_max, int* row_arg,
                        unsigned long long* col_key, float* row_mass, float* col_mass, int N0, int N1, int batch,
                        const int* n0s, const int* n1s, cudaStream_t st);
int launch_dual_softmax(const float* dist, long long d_bs, int ldd, const float* bin_score, float* P, long long p_bs,
                        int ldp, float* row_lse, float* col_lse, int N0, int N1, int batch, const int* n0s, const int* n1s,
                        cudaStream_t st);
long long sinkhorn_q_store_bytes(int batch, int N0max, int N1max, int storage);
int sinkhorn_rows_per_item(int batch, int N0max, int resident_ctas);
void sinkhorn_set_profiling(int on);
void sinkhorn_set_resident(int on);
void attention_set_variant(int v);
void gemm_set_variant(int v);
float sinkhorn_iter_ms();
int launch_sp_conv3x3(const imp_sp_conv_args& a, cudaStream_t st);
int launch_sp_conv1a(const float* img, const float* w, const float* bias, void* out_hi, void* out_lo, int B, int H, int W,
                     cudaStream_t st);
int launch_sp_maxpool2(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int B, int H, int W, int C, cudaStream_t st);
int launch_sp_scores(const float* logits, int ld, float* scores, int B, int Hc, int Wc, cudaStream_t st);
int launch_sp_nms(const float* scores, uint8_t* mask, uint8_t* supp, int B, int H, int W, int radius, cudaStream_t st);
int launch_sp_select(const imp_sp_select_args& a, cudaStream_t st);
int launch_sp_l2norm_rows(float* x, long long rows, int ld, cudaStream_t st);
int launch_sp_sample_descriptors(const float* dmap, const float* kpts_xy, const int* n_kpts, float* out, int Hc, int Wc, int max_k,
                                 cudaStream_t st);
}  // namespace imp

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

IMP_API const char* imp_last_error(void) { return imp::last_error(); }
IMP_API int imp_abi_version(void) { return IMP_B200_ABI_VERSION; }

IMP_API int imp_set_option(int32_t key, int32_t value) {
  switch (key) {
    case IMP_OPT_SK_RESIDENT: imp::sinkhorn_set_resident(value); return 0;
    case IMP_OPT_ATTN_VARIANT: imp::attention_set_variant(value); return 0;
    case IMP_OPT_GEMM_VARIANT: imp::gemm_set_variant(value); return 0;
    case IMP_OPT_SM_LIMIT: imp::set_sm_limit(value); return 0;
    default: imp::set_error("imp_set_option: unknown key %d", key); return 2;
  }
}

IMP_API int imp_split_planes(const float* x, const float* addend, void* hi, void* lo, int64_t n, void* stream) {
  return imp::launch_split_planes(x, addend, hi, lo, n, ST(stream));
}
IMP_API int imp_merge_planes(const void* hi, const void* lo, float* x, int64_t n, void* stream) {
  return imp::launch_merge_planes(hi, lo, x, n, ST(stream));
}
IMP_API int imp_gemm(const imp_gemm_args* args, void* stream) { return imp::launch_gemm(*args, ST(stream)); }
IMP_API int imp_attention(const imp_attn_args* args, void* stream) { return imp::launch_attention(*args, ST(stream)); }
IMP_API int imp_attention_colsum(const imp_attn_colsum_args* args, void* stream) {
  return imp::launch_attention_colsum(*args, ST(stream));
}
IMP_API int imp_instnorm_relu(const float* H, int64_t h_bs, int32_t ldh, const int32_t* ns, int32_t Nmax, int32_t C,
                      int32_t batch, float eps, int32_t relu, void* out_hi, void* out_lo, float* out_f32, int64_t o_bs,
                      int32_t ldo, void* stream) {
  return imp::launch_instnorm_relu_split(H, h_bs, ldh, ns, Nmax, C, batch, eps, relu, out_hi, out_lo, out_f32, o_bs, ldo,
                                         ST(stream));
}
IMP_API int imp_instnorm_apply(const float* H, const float* stat_partial, const float* stat_straddle, const int32_t* ns,
                       int32_t Np, int32_t C, int32_t images, float eps, int32_t relu, float* stats, void* out_hi,
                       void* out_lo, void* stream) {
  return imp::launch_instnorm_apply(H, stat_partial, stat_straddle, ns, Np, C, images, eps, relu, stats, out_hi, out_lo,
                                    ST(stream));
}
IMP_API int imp_kenc_input(const float* norm_kpts, const float* scores, float* out_xyz4, int64_t tokens, void* stream) {
  return imp::launch_kenc_input(norm_kpts, scores, out_xyz4, tokens, ST(stream));
}
IMP_API int imp_small_linear(const float* X, int32_t ldx, const float* W, const float* bias, float* Y, int32_t ldy, int64_t rows,
                     int32_t Cin, int32_t Cout, void* stream) {
  return imp::launch_small_linear(X, ldx, W, bias, Y, ldy, rows, Cin, Cout, ST(stream));
}
IMP_API int imp_sinkhorn(const imp_sinkhorn_args* args, void* stream) { return imp::launch_sinkhorn(*args, ST(stream)); }
IMP_API int64_t imp_sinkhorn_q_store_bytes(int32_t batch, int32_t N0max, int32_t N1max, int32_t storage) {
  return imp::sinkhorn_q_store_bytes(batch, N0max, N1max, storage);
}
IMP_API int imp_sinkhorn_rows_per_item(int32_t batch, int32_t N0max, int32_t resident_ctas) {
  return imp::sinkhorn_rows_per_item(batch, N0max, resident_ctas);
}
IMP_API int imp_set_profiling(int32_t on) {
  imp::sinkhorn_set_profiling(on);
  return 0;
}
IMP_API float imp_sinkhorn_iter_ms(void) { return imp::sinkhorn_iter_ms(); }
IMP_API int imp_matches(const imp_match_args* args, void* stream) { return imp::launch_matches(*args, ST(stream)); }
IMP_API int imp_dual_softmax(const float* dist, int64_t d_bs, int32_t ldd, const float* bin_score, float* P, int64_t p_bs,
                     int32_t ldp, float* row_lse, float* col_lse, int32_t N0, int32_t N1, int32_t batch, const int32_t* n0s,
                     const int32_t* n1s, void* stream) {
  return imp::launch_dual_softmax(dist, d_bs, ldd, bin_score, P, p_bs, ldp, row_lse, col_lse, N0, N1, batch, n0s, n1s,
                                  ST(stream));
}
IMP_API int imp_score_argmax(const float* P, int64_t p_bs, int32_t ldp, float* row_max, int32_t* row_arg, uint64_t* col_key,
                     float* row_mass, float* col_mass, int32_t N0, int32_t N1, int32_t batch, const int32_t* n0s,
                     const int32_t* n1s, void* stream) {
  return imp::launch_score_argmax(P, p_bs, ldp, row_max, row_arg, reinterpret_cast<unsigned long long*>(col_key),
                                  row_mass, col_mass, N0, N1, batch, n0s, n1s, ST(stream));
}
IMP_API int imp_pool_select(const imp_pool_args* args, void* stream) { return imp::launch_pool_select(*args, ST(stream)); }
IMP_API int imp_scatter_matches(const int64_t* idx0, const float* ms0, int32_t ld_sub, const int32_t* gids0, const int32_t* gids1,
                        int32_t ld_ids, const int32_t* cnt0, int64_t* out_idx, float* out_ms, int32_t ld_out, int32_t batch,
                        void* stream) {
  return imp::launch_scatter_matches(reinterpret_cast<const long long*>(idx0), ms0, ld_sub, gids0, gids1, ld_ids, cnt0,
                                     reinterpret_cast<long long*>(out_idx), out_ms, ld_out, batch, ST(stream));
}
IMP_API int imp_gather_rows(const void* in, int64_t in_bs, int32_t row_bytes_in, const int32_t* ids, int32_t ids_ld,
                    const int32_t* cnt, void* out, int64_t out_bs, int32_t row_bytes_out, int32_t copy_bytes,
                    int32_t max_rows, int32_t batch, void* stream) {
  return imp::launch_gather_rows(in, in_bs, row_bytes_in, ids, ids_ld, cnt, out, out_bs, row_bytes_out, copy_bytes,
                                 max_rows, batch, ST(stream));
}

IMP_API int imp_sp_conv3x3(const imp_sp_conv_args* args, void* stream) { return imp::launch_sp_conv3x3(*args, ST(stream)); }
IMP_API int imp_sp_conv1a(const float* img, const float* w, const float* bias, void* out_hi, void* out_lo, int32_t B, int32_t H,
                          int32_t W, void* stream) {
  return imp::launch_sp_conv1a(img, w, bias, out_hi, out_lo, B, H, W, ST(stream));
}
IMP_API int imp_sp_maxpool2(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int32_t B, int32_t H, int32_t W,
                            int32_t C, void* stream) {
  return imp::launch_sp_maxpool2(in_hi, in_lo, out_hi, out_lo, B, H, W, C, ST(stream));
}
IMP_API int imp_sp_scores(const float* logits, int32_t ld, float* scores, int32_t B, int32_t Hc, int32_t Wc, void* stream) {
  return imp::launch_sp_scores(logits, ld, scores, B, Hc, Wc, ST(stream));
}
IMP_API int imp_sp_nms(const float* scores, uint8_t* mask, uint8_t* supp, int32_t B, int32_t H, int32_t W, int32_t radius,
                       void* stream) {
  return imp::launch_sp_nms(scores, mask, supp, B, H, W, radius, ST(stream));
}
IMP_API int imp_sp_select(const imp_sp_select_args* args, void* stream) { return imp::launch_sp_select(*args, ST(stream)); }
IMP_API int imp_sp_l2norm_rows(float* x, int64_t rows, int32_t ld, void* stream) {
  return imp::launch_sp_l2norm_rows(x, rows, ld, ST(stream));
}
IMP_API int imp_sp_sample_descriptors(const float* dmap, const float* kpts_xy, const int32_t* n_kpts, float* out, int32_t Hc,
                                      int32_t Wc, int32_t max_k, void* stream) {
  return imp::launch_sp_sample_descriptors(dmap, kpts_xy, n_kpts, out, Hc, Wc, max_k, ST(stream));
}
}  // extern "C"
