// AudioResampler.h -- drop-in for glare::AudioResampler (audio/AudioResampler.h:23-42 in the reference tree): same
// class name, methods and argument meaning; the interpolation runs on the GPU through the C ABI
// (sbm_resampler_*, include/sbm_audio.h).  One instance = one stream, exactly like the reference class; callers with
// many streams should use the batched C ABI directly (one launch for all of them).
#pragma once
#ifndef SBM_USE_GLARE_CORE
#include "glare_core_min.h"
#endif
#include "../../include/sbm_audio.h"

namespace glare {

class AudioResampler {
public:
	AudioResampler();
	~AudioResampler();
	AudioResampler(const AudioResampler&) = delete;
	AudioResampler& operator=(const AudioResampler&) = delete;

	// AudioResampler::init (AudioResampler.cpp:24-36).  `device`: CUDA ordinal (the reference has no such argument).
	void init(int src_rate, int dest_rate, int device = 0);

	// AudioResampler::numSrcSamplesNeeded (AudioResampler.cpp:77-86).
	size_t numSrcSamplesNeeded(size_t dest_num_samples);

	// AudioResampler::resample (AudioResampler.cpp:89-113).  temp_buf is unused (kept for signature compatibility).
	// Throws glare::Exception if src_samples_size differs from numSrcSamplesNeeded(dest_samples_size) or CUDA fails.
	void resample(float* dest_samples, size_t dest_samples_size, const float* src_samples, size_t src_samples_size, js::Vector<float, 16>& temp_buf);

private:
	void ensure();
	sbm_resampler* r;
	int src_rate, dest_rate, device;
};

}  // namespace glare
