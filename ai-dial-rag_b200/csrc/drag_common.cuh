// Shared helpers for the drag_b200 C-ABI library (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/drag_b200.h"

namespace drag {

// thread-local last error message (drag_last_error)
char* error_buffer();
int fail(int code, const char* fmt, ...);

#define DRAG_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      return ::drag::fail(DRAG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                \
                          cudaGetErrorString(_e), __FILE__, __LINE__);                  \
  } while (0)

#define DRAG_REQUIRE(cond, ...)                                                         \
  do {                                                                                  \
    if (!(cond)) return ::drag::fail(DRAG_ERR_INVALID, __VA_ARGS__);                    \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int sm_count(int device);

// bf16 row-major [rows, cols] -> TMA map with a (box_rows x 64 columns) box and the 128-byte swizzle
// (out-of-bounds rows/columns read as zero and are clipped on stores).  The driver entry point
// cuTensorMapEncodeTiled is resolved at run time, so libcuda is not a link-time dependency.
int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);
// same with a box of box_cols (16, 32 or 64) columns: the swizzle span equals the box row (32/64/128 bytes)
int make_tmap_bf16_box(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols);

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

}  // namespace drag
