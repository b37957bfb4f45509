// device_internal.h -- shared by the CUDA translation units of libsubstrata_b200.so (not part of any interface).
#pragma once
#include <cuda_runtime.h>
#include "device_api.h"

namespace sbm {
namespace dev {
void note_launch(unsigned n = 1);            // bumps the library's kernel-launch counter
int check(cudaError_t e, const char* what);  // records the error text; returns 0 on success, -1 otherwise
int check_launch(const char* what);          // cudaGetLastError() after a launch
}  // namespace dev
}  // namespace sbm
