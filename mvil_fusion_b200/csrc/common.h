// common.h — error reporting and device gating shared by the translation units of libvils_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../include/vils_cabi.h"

namespace vils {
std::string& last_error();
int fail(int code, const std::string& msg);
int fail_cuda(cudaError_t e, const char* what);
// VILS_OK when `device` exists and is an sm_100-class part; there is no CPU fallback.
int require_device(int device);
}  // namespace vils
