// Error reporting and device helpers shared by the drag_b200 C-ABI library.
#include "drag_common.cuh"

#include <stdarg.h>

#include <mutex>

namespace drag {

char* error_buffer() {
  static thread_local char buf[1024] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 1024, fmt, ap);
  va_end(ap);
  return code;
}

int sm_count(int device) {
  static int cached[64] = {0};
  if (device < 0 || device >= 64) return -1;
  if (cached[device] > 0) return cached[device];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  cached[device] = n;
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  return make_tmap_bf16_box(map, base, rows, cols, box_rows, 64u);
}

int make_tmap_bf16_box(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
  if (box_cols != 16 && box_cols != 32 && box_cols != 64) return fail(DRAG_ERR_INVALID, "make_tmap_bf16_box: box_cols must be 16, 32 or 64");
  const CUtensorMapSwizzle swz = box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(DRAG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DRAG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return DRAG_OK;
}

}  // namespace drag

extern "C" const char* drag_last_error(void) { return drag::error_buffer(); }

extern "C" int drag_abi_version(void) { return DRAG_ABI_VERSION; }

extern "C" int drag_device_info(int device, int* n_devices, int* compute_capability, int* sm_count_out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (n_devices) *n_devices = 0;
    return drag::fail(DRAG_ERR_DEVICE, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
  }
  if (n_devices) *n_devices = n;
  if (device < 0 || device >= n) return drag::fail(DRAG_ERR_DEVICE, "device %d not present (%d devices)", device, n);
  int major = 0, minor = 0;
  DRAG_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  DRAG_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (compute_capability) *compute_capability = major * 10 + minor;
  if (sm_count_out) *sm_count_out = drag::sm_count(device);
  return DRAG_OK;
}
