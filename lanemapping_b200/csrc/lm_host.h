// lm_host.h -- host-side helpers shared by the translation units of liblm_bev.so (not exported API:
// the symbols are hidden from the C-ABI by name convention only; include/*.h is the boundary).
#pragma once
#include <cuda_runtime.h>

int lm_fail_msg(int code, const char *msg);          // sets the thread's lm_bev_last_error() text, returns code
int lm_cuda_fail(cudaError_t e, const char *what);   // same for a CUDA error, returns (int)e
int lm_sm_count();
