// Error reporting and device helpers shared by the drag_b200 C-ABI library.
#include "drag_common.cuh"

#include <stdarg.h>

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
