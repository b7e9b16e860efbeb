// Host-side helpers shared by the kernel launchers: error capture, driver entry points, tensor maps.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace imp {

void set_error(const char* fmt, ...);
const char* last_error();

#define IMP_CUDA_OK(expr)                                                                     \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      imp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)

#define IMP_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      imp::set_error(__VA_ARGS__);    \
      return 2;                       \
    }                                 \
  } while (0)

// 3-D fp16 tensor map {inner = k elements, rows, batch}, 128- or 64-byte swizzle, zero OOB fill.
// row_stride / batch_stride in elements.  box = {box_inner (= swizzle span: 64 or 32 elements), box_rows, 1}.
int make_tmap_f16_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batch,
                     uint64_t row_stride, uint64_t batch_stride, uint32_t box_inner, uint32_t box_rows,
                     int swizzle_bytes = 128);

// 3-D fp32 tensor map {inner, rows, batch}, box = {32 floats = one 128-byte swizzle span, box_rows, 1}
int make_tmap_f32_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batch, uint64_t row_stride,
                     uint64_t batch_stride, uint32_t box_rows);

// SMs the next launches may use: the device's count, or the limit the host layer set before launching onto a stream of an
// SM partition (green context) -- persistent kernels size their grids with it.
int num_sms();
void set_sm_limit(int n);

// cudaFuncSetAttribute state (dynamic shared memory limit, carve-out) is PER DEVICE: a launcher configures its kernels the
// first time it runs on each device of the process, not once per process.
struct DeviceOnce {
  unsigned long long mask = 0;
  bool first() {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask & bit) return false;
    mask |= bit;
    return true;
  }
};

}  // namespace imp
