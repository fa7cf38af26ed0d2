#include "common.h"

#include <atomic>

namespace vils {
std::string& last_error() { static thread_local std::string s; return s; }
int fail(int code, const std::string& msg) { last_error() = msg; return code; }
int fail_cuda(cudaError_t e, const char* what) {
  last_error() = std::string(what) + ": " + cudaGetErrorString(e);
  cudaGetLastError();
  return VILS_ERR_CUDA;
}
int require_device(int device) {
  // cudaGetDeviceProperties costs ~2 ms per call: the verdict for a device ordinal is cached (stateless entry points call this every time)
  static std::atomic<int> verdict[64];
  if (device >= 0 && device < 64 && verdict[device].load(std::memory_order_acquire) == 1)
    return cudaSetDevice(device) == cudaSuccess ? VILS_OK : fail(VILS_ERR_CUDA, "cudaSetDevice failed");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(VILS_ERR_NO_DEVICE, "no CUDA device (libvils_b200 has no CPU path)"); }
  if (device < 0 || device >= n) return fail(VILS_ERR_NO_DEVICE, "device ordinal out of range");
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) return fail(VILS_ERR_NO_DEVICE, "cudaDeviceGetAttribute failed");
  if (major != 10) return fail(VILS_ERR_NO_DEVICE, "libvils_b200 is built for sm_100a only");
  if (cudaSetDevice(device) != cudaSuccess) return fail(VILS_ERR_CUDA, "cudaSetDevice failed");
  if (device < 64) verdict[device].store(1, std::memory_order_release);
  return VILS_OK;
}
}  // namespace vils
