// AudioResampler.cpp -- see AudioResampler.h.  Host logic only: all sample arithmetic is in k5_resampler.cu.
#include "AudioResampler.h"

namespace glare {

AudioResampler::AudioResampler() : r(nullptr), src_rate(0), dest_rate(0) {}

AudioResampler::~AudioResampler() { if (r) sbm_resampler_destroy(r); }

void AudioResampler::init(int src_rate_, int dest_rate_, int device) {
	if (r) { sbm_resampler_destroy(r); r = nullptr; }
	src_rate = src_rate_; dest_rate = dest_rate_;
	const int rc = sbm_resampler_create(device, 1, &src_rate, dest_rate, &r);
	if (rc != SBM_OK) throw glare::Exception("AudioResampler::init: sbm_resampler_create failed (" + std::to_string(rc) + ")");
}

size_t AudioResampler::numSrcSamplesNeeded(size_t dest_num_samples) {
	if (!r) throw glare::Exception("AudioResampler::numSrcSamplesNeeded: init() has not been called");
	size_t n = 0;
	sbm_resampler_num_src_needed(r, dest_num_samples, &n);
	return n;
}

void AudioResampler::resample(float* dest_samples, size_t dest_samples_size, const float* src_samples, size_t src_samples_size, js::Vector<float, 16>& /*temp_buf*/) {
	if (!r) throw glare::Exception("AudioResampler::resample: init() has not been called");
	if (src_samples_size != numSrcSamplesNeeded(dest_samples_size))
		throw glare::Exception("AudioResampler::resample: src_samples_size must equal numSrcSamplesNeeded(dest_samples_size)");
	const size_t off = 0;
	const int rc = sbm_resampler_resample(r, dest_samples, dest_samples_size, src_samples, &off, src_samples_size);
	if (rc != SBM_OK) throw glare::Exception("AudioResampler::resample: sbm_resampler_resample failed (" + std::to_string(rc) + ")");
}

}  // namespace glare
