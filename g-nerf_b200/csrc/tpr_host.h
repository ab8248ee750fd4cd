// Host-side helpers shared by the translation units of libtriplane_b200.so (defined in triplane_b200.cu).
#pragma once
#include <cuda_runtime.h>

namespace tpr {

// set tpr_last_error()'s thread-local message and return `code` (argument errors, < 0) / the CUDA error (> 0)
int fail(int code, const char* msg);
int cuda_fail(cudaError_t e, const char* what);

struct DeviceInfo { int sms = 0; int smem_optin = 0; bool ok = false; };
DeviceInfo device_info();        // of the current device, cached

// blocks for a grid-stride kernel: enough for the work, at most `waves` per SM
int grid_for(long long work_items, int per_block, int sms, int waves);

}  // namespace tpr
