#include "common.h"

#include <stdarg.h>
#include <string.h>

#include <mutex>

namespace imp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is not a link-time dependency (the build box has no driver): resolve through the runtime.
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_f16_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batch,
                     uint64_t row_stride, uint64_t batch_stride, uint32_t box_inner, uint32_t box_rows,
                     int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  IMP_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  IMP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base not 16-byte aligned");
  IMP_REQUIRE((row_stride * 2) % 16 == 0 && (batch_stride * 2) % 16 == 0, "tensor map strides must be 16-byte multiples");
  cuuint64_t gdim[3] = {inner, rows, batch};
  cuuint64_t gstr[2] = {row_stride * 2, batch_stride * 2};
  if (batch == 1 && gstr[1] == 0) gstr[1] = gstr[0] * rows;
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IMP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): dims %llu x %llu x %llu stride %llu/%llu box %u x %u",
              (int)r, (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)batch,
              (unsigned long long)row_stride, (unsigned long long)batch_stride, box_inner, box_rows);
  return 0;
}

int make_tmap_f32_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batch, uint64_t row_stride,
                     uint64_t batch_stride, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  IMP_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  IMP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (row_stride * 4) % 16 == 0, "fp32 tensor map: 16-byte alignment");
  cuuint64_t gdim[3] = {inner, rows, batch};
  cuuint64_t gstr[2] = {row_stride * 4, batch_stride * 4};
  if (batch == 1 && gstr[1] == 0) gstr[1] = gstr[0] * rows;
  cuuint32_t box[3] = {32, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IMP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp32) failed (%d)", (int)r);
  return 0;
}

static int g_sm_limit = 0;
void set_sm_limit(int n) { g_sm_limit = n > 0 ? n : 0; }

int num_sms() {
  if (g_sm_limit > 0) return g_sm_limit;
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace imp
