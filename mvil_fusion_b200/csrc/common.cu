#include "common.h"

namespace vils {
std::string& last_error() { static thread_local std::string s; return s; }
int fail(int code, const std::string& msg) { last_error() = msg; return code; }
int fail_cuda(cudaError_t e, const char* what) {
  last_error() = std::string(what) + ": " + cudaGetErrorString(e);
  cudaGetLastError();
  return VILS_ERR_CUDA;
}
int require_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(VILS_ERR_NO_DEVICE, "no CUDA device (libvils_b200 has no CPU path)"); }
  if (device < 0 || device >= n) return fail(VILS_ERR_NO_DEVICE, "device ordinal out of range");
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return fail(VILS_ERR_NO_DEVICE, "cudaGetDeviceProperties failed");
  if (p.major != 10) return fail(VILS_ERR_NO_DEVICE, "libvils_b200 is built for sm_100a only");
  if (cudaSetDevice(device) != cudaSuccess) return fail(VILS_ERR_CUDA, "cudaSetDevice failed");
  return VILS_OK;
}
}  // namespace vils
