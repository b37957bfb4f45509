// device_internal.h -- shared by the CUDA translation units of libsubstrata_b200.so (not part of any interface).
#pragma once
#include <cuda_runtime.h>
#include "device_api.h"

#ifndef SBM_PDL
#define SBM_PDL 1
#endif

namespace sbm {
namespace dev {
void note_launch(unsigned n = 1);            // bumps the library's kernel-launch counter
int check(cudaError_t e, const char* what);  // records the error text; returns 0 on success, -1 otherwise
int check_launch(const char* what);          // cudaGetLastError() after a launch
int current_sm_count();                      // SM count of the current device (cached per device, k2_mix.cu)
}  // namespace dev

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl may be set up -- and start running its
// prologue -- while the previous kernel of the stream is still draining; it MUST call grid_dep_wait() before it touches
// anything that kernel wrote.  (Without the launch attribute the wait is a no-op.)  The per-buffer chain K1 -> K2 -> K3 ->
// [gather] -> K4a -> K4b is five short-lived launches; this takes their launch latencies off the critical path.
#if defined(__CUDACC__)
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// The primary's side: "the next kernel of the stream may be set up now".  Its CTAs become resident as soon as EVERY CTA of this grid
// has executed this (or exited) and resources allow.
__device__ __forceinline__ void grid_dep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = SBM_PDL;
	cfg.attrs = attr; cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif
}  // namespace sbm
