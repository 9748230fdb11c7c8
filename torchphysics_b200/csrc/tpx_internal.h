// Shared by the translation units of libtpx.so: per-thread error string helpers.
#pragma once
#include <cuda_runtime.h>

namespace tpx {
int fail(int code, const char* fmt, ...);      // records the message, returns code
int cuda_fail(const char* what);               // TPX_E_CUDA with cudaGetLastError() text
const char* last_error();
}  // namespace tpx
