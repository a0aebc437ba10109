// Shared host-side plumbing of the C-ABI library: error slot, CUDA check macros, small helpers.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/millieye_b200.h"

namespace me {

// Last error text of the calling thread; read back through me_last_error().
char* error_slot();
int fail(int code, const char* fmt, ...);

#define ME_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::me::fail(ME_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define ME_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) return ::me::fail(ME_ERR_ARG, __VA_ARGS__);    \
  } while (0)

#define ME_LAUNCH_CHECK() ME_CUDA(cudaGetLastError())

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

int sm_count();   // of the current device (cached per device ordinal)
// true the first time it is called with `seen` on the current device: per-device one-time setup such as
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize), which belongs to the device's copy of the function
bool first_use_on_device(bool (&seen)[64]);

// ---- tensor-map encoders (driver entry points fetched at run time; no link-time libcuda) ----
// 2D row-major [rows][cols] view with `pitch_elems` elements between rows.
int encode_tiled_2d(CUtensorMap* map, CUtensorMapDataType dt, int esize, const void* base, uint64_t cols,
                    uint64_t rows, uint64_t pitch_elems, uint32_t box_cols, uint32_t box_rows,
                    CUtensorMapSwizzle swz);
// NHWC fp16 activation viewed as (C, W, H, N) for im2col-mode loads of a KxK / pad / stride conv.
int encode_im2col_nhwc(CUtensorMap* map, const void* base, int n, int h, int w, int c, int pitch_elems,
                       int ksize, int pad, int stride, uint32_t channels_per_pixel, uint32_t pixels_per_col,
                       CUtensorMapSwizzle swz);

static inline CUtensorMapSwizzle swizzle_for_row_bytes(int row_bytes) {
  switch (row_bytes) {
    case 128: return CU_TENSOR_MAP_SWIZZLE_128B;
    case 64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case 32: return CU_TENSOR_MAP_SWIZZLE_32B;
    default: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
}

}  // namespace me
