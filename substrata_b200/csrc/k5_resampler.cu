// k5_resampler.cu -- glare::AudioResampler drop-in (audio/AudioResampler.h:23-42) for callers that resample outside
// the engine: `count` independent streams, each converting src_rate -> dst_rate by the reference's linear
// interpolation with state carried across calls (AR.cpp:24-113).  Host pointers in, host pointers out; the
// arithmetic is the same resampled_sample() the mix kernel uses (k2_core.h), one thread per output sample.
#include <string.h>

#include <new>
#include <vector>

#include "../../include/sbm_audio.h"
#include "device_internal.h"
#include "k2_core.h"

using namespace sbm;

namespace {

__global__ void __launch_bounds__(256) k5_resample_kernel(const WorkRec* __restrict__ recs, float* __restrict__ dst, uint32_t n_dst,
                                                         float2* __restrict__ tails) {
	const WorkRec r = recs[blockIdx.y];
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_dst) dst[(size_t)blockIdx.y * n_dst + i] = resampled_sample(r, i);
	if (i == 0) {
		float p0, p1;
		resampler_tail(r, &p0, &p1);
		tails[blockIdx.y] = make_float2(p0, p1);
	}
}

template <class T> bool grow(T*& p, size_t& cap, size_t need) {
	if (need <= cap) return true;
	if (p) cudaFree(p);
	p = nullptr; cap = 0;
	if (cudaMalloc((void**)&p, sizeof(T) * need) != cudaSuccess) return false;
	cap = need;
	return true;
}

}  // namespace

struct sbm_resampler {
	int device = 0;
	cudaStream_t stream = nullptr;
	int count = 0;
	int dst_rate = 0;
	std::vector<int> src_rate;
	std::vector<Resamp> state;
	std::vector<WorkRec> recs;
	std::vector<float2> tails;
	WorkRec* d_recs = nullptr; size_t recs_cap = 0;
	float2* d_tails = nullptr; size_t tails_cap = 0;
	float* d_src = nullptr; size_t src_cap = 0;
	float* d_dst = nullptr; size_t dst_cap = 0;
};

namespace {
// AudioResampler::numSrcSamplesNeeded (AR.cpp:77-86)
size_t needed_for(const sbm_resampler* r, int k, size_t n_dst) {
	const Resamp& s = r->state[(size_t)k];
	const int64_t max_dest = s.prev_dest_dst_coords + (int64_t)n_dst;
	const double max_dest_src = (double)max_dest * (double)r->src_rate[(size_t)k] / (double)r->dst_rate;
	const int64_t prev1 = s.prev_samples_0_src_coords + 1;
	return (size_t)ceil(max_dest_src) - (size_t)prev1;
}
}  // namespace

extern "C" {

SBM_API int sbm_resampler_create(int device, int count, const int* src_rate, int dst_rate, sbm_resampler** out) {
	if (!out || count <= 0 || !src_rate || dst_rate <= 0) return SBM_ERR_INVALID_ARG;
	*out = nullptr;
	for (int k = 0; k < count; ++k) if (src_rate[k] <= 0) return SBM_ERR_INVALID_ARG;
	void* st = nullptr;
	if (dev::init(device, &st)) return SBM_ERR_CUDA;
	sbm_resampler* r = new (std::nothrow) sbm_resampler();
	if (!r) return SBM_ERR_NO_MEMORY;
	r->device = device; r->stream = (cudaStream_t)st; r->count = count; r->dst_rate = dst_rate;
	r->src_rate.assign(src_rate, src_rate + count);
	Resamp init; init.prev_dest_dst_coords = -1; init.prev_samples_0_src_coords = -2; init.prev_samples[0] = 0.0f; init.prev_samples[1] = 0.0f;   // AR.cpp:24-36
	r->state.assign((size_t)count, init);
	r->recs.resize((size_t)count);
	r->tails.resize((size_t)count);
	*out = r;
	return SBM_OK;
}

SBM_API void sbm_resampler_destroy(sbm_resampler* r) {
	if (!r) return;
	cudaSetDevice(r->device);
	if (r->stream) { cudaStreamSynchronize(r->stream); cudaStreamDestroy(r->stream); }
	cudaFree(r->d_recs); cudaFree(r->d_tails); cudaFree(r->d_src); cudaFree(r->d_dst);
	delete r;
}

SBM_API int sbm_resampler_num_src_needed(sbm_resampler* r, size_t n_dst, size_t* out_needed) {
	if (!r || !out_needed) return SBM_ERR_INVALID_ARG;
	for (int k = 0; k < r->count; ++k) out_needed[k] = needed_for(r, k, n_dst);
	return SBM_OK;
}

SBM_API int sbm_resampler_resample(sbm_resampler* r, float* dst, size_t n_dst, const float* src, const size_t* src_offsets, size_t src_total) {
	if (!r || !dst || !src_offsets || n_dst == 0 || n_dst > 0x7fffffffu || (!src && src_total)) return SBM_ERR_INVALID_ARG;
	if (dev::use(r->device)) return SBM_ERR_CUDA;
	const size_t count = (size_t)r->count;
	if (!grow(r->d_recs, r->recs_cap, count) || !grow(r->d_tails, r->tails_cap, count) || !grow(r->d_src, r->src_cap, src_total + 8) ||
	    !grow(r->d_dst, r->dst_cap, count * n_dst))
		return SBM_ERR_NO_MEMORY;
	for (size_t k = 0; k < count; ++k) {
		const size_t need = needed_for(r, (int)k, n_dst);
		if (src_offsets[k] + need > src_total) return SBM_ERR_INVALID_ARG;
		const Resamp& s = r->state[k];
		WorkRec& w = r->recs[k];
		memset(&w, 0, sizeof(w));
		w.pcm = r->d_src + src_offsets[k]; w.len = (uint32_t)need; w.start = 0; w.n_src = (uint32_t)need; w.avail = (uint32_t)need;
		w.flags = kRecResample;
		// AudioResampler::resample coordinates (AR.cpp:101-103)
		const int64_t dest0 = s.prev_dest_dst_coords + 1;
		const double x0 = (double)dest0 * (double)r->src_rate[k] / (double)r->dst_rate - (double)s.prev_samples_0_src_coords;
		w.x0_fx = (unsigned long long)((x0 > 0.0 ? x0 : 0.0) * kFx);
		w.step_fx = (unsigned long long)((double)r->src_rate[k] / (double)r->dst_rate * kFx);
		w.prev0 = s.prev_samples[0]; w.prev1 = s.prev_samples[1];
	}
	cudaStream_t st = r->stream;
	if (dev::check(cudaMemcpyAsync(r->d_recs, r->recs.data(), sizeof(WorkRec) * count, cudaMemcpyHostToDevice, st), "resampler H2D") ||
	    (src_total && dev::check(cudaMemcpyAsync(r->d_src, src, sizeof(float) * src_total, cudaMemcpyHostToDevice, st), "resampler H2D")))
		return SBM_ERR_CUDA;
	const dim3 grid((unsigned)((n_dst + 255) / 256), (unsigned)count);
	k5_resample_kernel<<<grid, 256, 0, st>>>(r->d_recs, r->d_dst, (uint32_t)n_dst, r->d_tails);
	dev::note_launch();
	if (dev::check_launch("k5_resample") ||
	    dev::check(cudaMemcpyAsync(dst, r->d_dst, sizeof(float) * count * n_dst, cudaMemcpyDeviceToHost, st), "resampler D2H") ||
	    dev::check(cudaMemcpyAsync(r->tails.data(), r->d_tails, sizeof(float2) * count, cudaMemcpyDeviceToHost, st), "resampler D2H") ||
	    dev::check(cudaStreamSynchronize(st), "resampler sync"))
		return SBM_ERR_CUDA;
	for (size_t k = 0; k < count; ++k) {
		Resamp& s = r->state[k];
		s.prev_dest_dst_coords += (int64_t)n_dst;
		s.prev_samples_0_src_coords += (int64_t)r->recs[k].n_src;
		s.prev_samples[0] = r->tails[k].x; s.prev_samples[1] = r->tails[k].y;
	}
	return SBM_OK;
}

}  // extern "C"
