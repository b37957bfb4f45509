// AudioResampler.cpp -- see AudioResampler.h.  Host logic only: all sample arithmetic is in k5_resampler.cu.
#include "AudioResampler.h"

namespace glare {

AudioResampler::AudioResampler() : r(nullptr), src_rate(0), dest_rate(0), device(0) {}

AudioResampler::~AudioResampler() { if (r) sbm_resampler_destroy(r); }

// init only records the rates (and forgets the carried state, AudioResampler.cpp:24-36): every AudioSource owns an AudioResampler
// (AudioEngine.h:153) that the accelerated engine never calls -- sources are resampled inside the mix kernels -- so the device
// object is created by the first call that needs it.
void AudioResampler::init(int src_rate_, int dest_rate_, int device_) {
	if (r) { sbm_resampler_destroy(r); r = nullptr; }
	src_rate = src_rate_; dest_rate = dest_rate_; device = device_;
}

void AudioResampler::ensure() {
	if (r) return;
	if (src_rate <= 0 || dest_rate <= 0) throw glare::Exception("AudioResampler: init() has not been called");
	const int rc = sbm_resampler_create(device, 1, &src_rate, dest_rate, &r);
	if (rc != SBM_OK) throw glare::Exception("AudioResampler: sbm_resampler_create failed (" + std::to_string(rc) + ")");
}

size_t AudioResampler::numSrcSamplesNeeded(size_t dest_num_samples) {
	ensure();
	size_t n = 0;
	sbm_resampler_num_src_needed(r, dest_num_samples, &n);
	return n;
}

void AudioResampler::resample(float* dest_samples, size_t dest_samples_size, const float* src_samples, size_t src_samples_size, js::Vector<float, 16>& /*temp_buf*/) {
	ensure();
	if (src_samples_size != numSrcSamplesNeeded(dest_samples_size))
		throw glare::Exception("AudioResampler::resample: src_samples_size must equal numSrcSamplesNeeded(dest_samples_size)");
	const size_t off = 0;
	const int rc = sbm_resampler_resample(r, dest_samples, dest_samples_size, src_samples, &off, src_samples_size);
	if (rc != SBM_OK) throw glare::Exception("AudioResampler::resample: sbm_resampler_resample failed (" + std::to_string(rc) + ")");
}

}  // namespace glare
