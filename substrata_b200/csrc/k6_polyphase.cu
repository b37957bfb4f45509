// k6_polyphase.cu -- "K6": the arithmetic of Resonance's polyphase resampler (RA/dsp/resampler.cc:63-135).  One thread per
// (channel, output sample): output k reads the coeffs_per_phase input samples that end at
//   input_sample(k) = last + (time_modulo + k * down) / up      with the filter phase  (time_modulo + k * down) % up,
// taken from {carried state, input} (polyphase.cc lays them out back to back) and sums the products in the reference's order --
// oldest sample first, separate multiply and add (no FMA contraction), starting from 0 -- so the result is the reference's bits.
// HBM-bound in principle (4 B in, 4 * up / down B out per sample); the coefficient table (<= 64 KB) stays in L1 / L2.
#include "device_internal.h"

namespace sbm {

__global__ void __launch_bounds__(256) k6_polyphase_kernel(PolyphaseArgs a) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
	if (k >= a.n_out) return;
	const unsigned long long total = (unsigned long long)a.time_modulo + (unsigned long long)k * a.down;
	const uint32_t input_sample = a.last + (uint32_t)(total / a.up), phase = (uint32_t)(total % a.up);
	// window: samples input_sample - cpp + 1 .. input_sample of the input = positions + hist in the extended array
	const float* __restrict__ x = a.in + (size_t)c * a.ext + (size_t)input_sample;      // (input_sample - cpp + 1) + hist == input_sample
	const float* __restrict__ h = a.coeffs + (size_t)phase * a.cpp;
	float acc = 0.0f;
	for (uint32_t j = 0; j < a.cpp; ++j) acc = __fadd_rn(acc, __fmul_rn(x[j], h[j]));
	a.out[(size_t)c * a.n_out + k] = acc;
}

namespace dev {
int launch_polyphase(const PolyphaseArgs& a, void* stream) {
	if (!a.n_out) return 0;
	k6_polyphase_kernel<<<dim3((a.n_out + 255) / 256, a.channels), 256, 0, (cudaStream_t)stream>>>(a);
	note_launch();
	return check_launch("k6_polyphase");
}
}  // namespace dev
}  // namespace sbm
