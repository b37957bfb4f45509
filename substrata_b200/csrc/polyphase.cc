// polyphase.cc -- drop-in for vraudio::Resampler, Resonance Audio's polyphase sample-rate converter (RA/dsp/resampler.h:27-130,
// RA/dsp/resampler.cc), and for the one place the AudioEngine path uses it: bringing the SH-HRIRs to a device rate other than
// 48 kHz (CreateShHrirsFromWav, RA/dsp/sh_hrir_creator.cc:42-55).  SURVEY.md 8f row 4.
//
// Host side: rate bookkeeping, the windowed-sinc prototype filter in polyphase order and the carried state -- float for float
// like the reference (built without FMA contraction; the same libm sinf / cosf) -- so that the coefficients are the
// reference's bits.  Device side (k6_polyphase.cu): every output sample is an independent dot product of coeffs_per_phase taps;
// products are summed in the reference's order with separate multiplies and adds, so the OUTPUT is bit-exact too.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <numeric>
#include <vector>

#include "../../include/sbm_audio.h"
#include "device_api.h"

using namespace sbm;

namespace {
constexpr size_t kMaxFrames16k = 16384;        // kMaxSupportedNumFrames (RA/base/constants_and_types.h)
constexpr size_t kMaxChannels = 16;            // (kMaxSupportedAmbisonicOrder + 1)^2
constexpr size_t kTransitionBandwidthRatio = 13;
size_t gcd(size_t a, size_t b) { while (b) { const size_t t = a % b; a = b; b = t; } return a; }
}  // namespace

struct sbm_polyphase {
	int device = 0;
	void* stream = nullptr;
	size_t up = 0, down = 0, time_modulo_up_rate = 0, last_processed_sample = 0, num_channels = 0, coeffs_per_phase = 0;
	size_t io_channels = 0;         // channels of the caller's buffers (the reference reads them off its AudioBuffers)
	std::vector<float> coeffs;      // transposed_filter_coeffs_ [kMaxFrames16k]
	std::vector<float> state;       // state_ [kMaxChannels][kMaxFrames16k]
	bool coeffs_dirty = true;
	float* d_coeffs = nullptr; float* d_in = nullptr; float* d_out = nullptr;
	size_t in_cap = 0, out_cap = 0;
	std::vector<float> h_in;
};

namespace {

// Resampler::GenerateInterpolatingFilter + GenerateSincFilter + ArrangeFilterAsPolyphase (resampler.cc:258-333), on a full Hann
// window (RA/dsp/utils.cc:173-188).
void generate_filter(sbm_polyphase* p, int sample_rate) {
	const size_t max_rate = std::max(p->up, p->down);
	const float cutoff_frequency = (float)sample_rate / (float)(2 * max_rate);
	size_t filter_length = max_rate * kTransitionBandwidthRatio;
	filter_length += filter_length % 2;
	std::vector<float> f(kMaxFrames16k, 0.0f);
	const float two_pi = 6.28318530717958647692f;
	const float angular_cutoff_frequency = two_pi * cutoff_frequency / (float)sample_rate;
	const size_t half_filter_length = filter_length / 2;
	const float scaling_factor = two_pi / ((float)filter_length - 1.0f);
	for (size_t i = 0; i < filter_length; ++i) f[i] = 0.5f * (1.0f - cosf(scaling_factor * (float)i));
	for (size_t i = 0; i < filter_length; ++i) {
		if (i == half_filter_length) f[half_filter_length] *= angular_cutoff_frequency;
		else {
			const float denominator = (float)i - ((float)filter_length / 2.0f);
			f[i] *= sinf(angular_cutoff_frequency * denominator) / denominator;
		}
	}
	const float normalizing_factor = (float)p->up / std::accumulate(f.begin(), f.begin() + (long)filter_length, 0.0f);
	for (size_t i = 0; i < filter_length; ++i) f[i] = normalizing_factor * f[i];
	// pad so that the filter divides into max_rate phases; coefficients transposed and flipped per phase
	const size_t transposed_length = filter_length + max_rate - (filter_length % max_rate);
	p->coeffs_per_phase = transposed_length / max_rate;
	std::fill(p->coeffs.begin(), p->coeffs.end(), 0.0f);
	for (size_t i = 0; i < p->up; ++i)
		for (size_t j = 0; j < p->coeffs_per_phase; ++j)
			if (j * p->up + i < filter_length) p->coeffs[(p->coeffs_per_phase - 1 - j) + i * p->coeffs_per_phase] = f[j * p->up + i];
	p->coeffs_dirty = true;
}

size_t max_output_length(const sbm_polyphase* p, size_t n) { return p->up == p->down ? n : (n * p->up) / p->down + 1; }
size_t next_output_length(const sbm_polyphase* p, size_t n) {
	if (p->up == p->down) return n;
	const size_t mx = max_output_length(p, n);
	return (p->time_modulo_up_rate + p->up * p->last_processed_sample) >= ((n * p->up) % p->down) ? mx - 1 : mx;
}

template <class T> bool grow(T*& ptr, size_t& cap, size_t need) {
	if (need <= cap) return true;
	if (ptr) dev::release(ptr);
	cap = std::max(need, cap * 2);
	ptr = (T*)dev::alloc(sizeof(T) * cap);
	if (!ptr) { cap = 0; return false; }
	return true;
}

}  // namespace

extern "C" {

SBM_API int sbm_polyphase_create(int device, sbm_polyphase** out) {
	if (!out) return SBM_ERR_INVALID_ARG;
	*out = nullptr;
	void* st = nullptr;
	if (dev::init(device, &st)) return SBM_ERR_CUDA;
	sbm_polyphase* p = new (std::nothrow) sbm_polyphase();
	if (!p) return SBM_ERR_NO_MEMORY;
	p->device = device; p->stream = st;
	p->coeffs.assign(kMaxFrames16k, 0.0f);
	p->state.assign(kMaxChannels * kMaxFrames16k, 0.0f);
	p->d_coeffs = (float*)dev::alloc(sizeof(float) * kMaxFrames16k);
	if (!p->d_coeffs) { dev::shutdown(st); delete p; return SBM_ERR_CUDA; }
	*out = p;
	return SBM_OK;
}

SBM_API void sbm_polyphase_destroy(sbm_polyphase* p) {
	if (!p) return;
	dev::use(p->device);
	if (p->stream) dev::sync(p->stream);
	dev::release(p->d_coeffs); dev::release(p->d_in); dev::release(p->d_out);
	dev::shutdown(p->stream);
	delete p;
}

// Resampler::AreSampleRatesSupported (resampler.cc:216-226).
SBM_API int sbm_polyphase_rates_supported(int source_hz, int destination_hz) {
	if (source_hz <= 0 || destination_hz <= 0) return 0;
	const size_t max_rate = (size_t)std::max(source_hz, destination_hz) / gcd((size_t)source_hz, (size_t)destination_hz);
	size_t filter_length = max_rate * kTransitionBandwidthRatio;
	filter_length += filter_length % 2;
	return filter_length <= kMaxFrames16k ? 1 : 0;
}

// Resampler::SetRateAndNumChannels (resampler.cc:176-214).
SBM_API int sbm_polyphase_set_rate_and_channels(sbm_polyphase* p, int source_hz, int destination_hz, int num_channels) {
	if (!p || source_hz <= 0 || destination_hz <= 0 || num_channels <= 0 || (size_t)num_channels > kMaxChannels) return SBM_ERR_INVALID_ARG;
	if (!sbm_polyphase_rates_supported(source_hz, destination_hz)) return SBM_ERR_UNSUPPORTED;
	const size_t g = gcd((size_t)destination_hz, (size_t)source_hz);
	const size_t destination = (size_t)destination_hz / g, source = (size_t)source_hz / g;
	const size_t old_state_size = p->coeffs_per_phase > 0 ? p->coeffs_per_phase - 1 : 0;
	p->io_channels = (size_t)num_channels;
	if (destination != p->up || source != p->down) {
		p->up = destination; p->down = source;
		if (p->up == p->down) return SBM_OK;
		generate_filter(p, source_hz);
		p->time_modulo_up_rate = 0;
	}
	if (p->num_channels != (size_t)num_channels) {
		p->num_channels = (size_t)num_channels;
		if (p->up != p->down) {      // InitializeStateBuffer (resampler.cc:236-256)
			const size_t new_state_size = p->coeffs_per_phase > 0 ? p->coeffs_per_phase - 1 : 0;
			if (old_state_size != new_state_size) {
				const size_t lo = std::min(new_state_size, old_state_size), hi = std::max(new_state_size, old_state_size);
				for (size_t c = 0; c < p->num_channels; ++c) std::fill(p->state.begin() + (long)(c * kMaxFrames16k + lo), p->state.begin() + (long)(c * kMaxFrames16k + hi), 0.0f);
			}
		}
	}
	return SBM_OK;
}

SBM_API void sbm_polyphase_reset_state(sbm_polyphase* p) {      // Resampler::ResetState (resampler.cc:228-234)
	if (!p) return;
	p->time_modulo_up_rate = 0; p->last_processed_sample = 0;
	std::fill(p->state.begin(), p->state.end(), 0.0f);
}

SBM_API size_t sbm_polyphase_max_output_length(const sbm_polyphase* p, size_t input_length) { return p ? max_output_length(p, input_length) : 0; }
SBM_API size_t sbm_polyphase_next_output_length(const sbm_polyphase* p, size_t input_length) { return p ? next_output_length(p, input_length) : 0; }

// Resampler::Process (resampler.cc:63-160).  input: planar [num_channels][n_in]; output: planar [num_channels][n_out] with
// n_out == sbm_polyphase_next_output_length(n_in).
SBM_API int sbm_polyphase_process(sbm_polyphase* p, const float* input, size_t n_in, float* output, size_t n_out) {
	if (!p || !input || !output || !p->io_channels || n_in == 0) return SBM_ERR_INVALID_ARG;
	if (n_out != next_output_length(p, n_in)) return SBM_ERR_INVALID_ARG;
	const size_t nch = p->io_channels;
	if (p->up == p->down) { memcpy(output, input, sizeof(float) * nch * n_in); return SBM_OK; }
	if (dev::use(p->device)) return SBM_ERR_CUDA;
	const size_t cpp = p->coeffs_per_phase, hist = cpp - 1;
	// the kernel's input: per channel, the carried state followed by this call's samples
	const size_t ext = hist + n_in;
	p->h_in.resize(nch * ext);
	for (size_t c = 0; c < nch; ++c) {
		memcpy(p->h_in.data() + c * ext, p->state.data() + c * kMaxFrames16k, sizeof(float) * hist);
		memcpy(p->h_in.data() + c * ext + hist, input + c * n_in, sizeof(float) * n_in);
	}
	if (!grow(p->d_in, p->in_cap, nch * ext) || !grow(p->d_out, p->out_cap, nch * std::max<size_t>(n_out, 1))) return SBM_ERR_NO_MEMORY;
	if (p->coeffs_dirty) {
		if (dev::h2d(p->d_coeffs, p->coeffs.data(), sizeof(float) * kMaxFrames16k, p->stream)) return SBM_ERR_CUDA;
		p->coeffs_dirty = false;
	}
	if (n_out) {
		PolyphaseArgs a;
		a.in = p->d_in; a.coeffs = p->d_coeffs; a.out = p->d_out; a.channels = (uint32_t)nch; a.ext = (uint32_t)ext; a.hist = (uint32_t)hist; a.n_out = (uint32_t)n_out;
		a.up = (uint32_t)p->up; a.down = (uint32_t)p->down; a.time_modulo = (uint32_t)p->time_modulo_up_rate; a.last = (uint32_t)p->last_processed_sample; a.cpp = (uint32_t)cpp;
		if (dev::h2d(p->d_in, p->h_in.data(), sizeof(float) * nch * ext, p->stream) || dev::launch_polyphase(a, p->stream) ||
		    dev::d2h(output, p->d_out, sizeof(float) * nch * n_out, p->stream) || dev::sync(p->stream))
			return SBM_ERR_CUDA;
	}
	// where the read position and the filter phase end up: n_out steps of  t += down; input_sample += t / up; t %= up
	const unsigned long long total = (unsigned long long)p->time_modulo_up_rate + (unsigned long long)n_out * p->down;
	const size_t input_sample = p->last_processed_sample + (size_t)(total / p->up);
	p->time_modulo_up_rate = (size_t)(total % p->up);
	p->last_processed_sample = input_sample - n_in;
	// the |state_| buffer, exactly as the reference maintains it (its channels are kMaxSupportedNumFrames long)
	const long left = (long)cpp - 1 - (long)n_in;
	for (size_t c = 0; c < nch; ++c) {
		float* st = p->state.data() + c * kMaxFrames16k;
		if (left > 0) {
			memmove(st, st + kMaxFrames16k - (size_t)left, sizeof(float) * (size_t)left);
			memcpy(st + kMaxFrames16k - n_in, input + c * n_in, sizeof(float) * n_in);
		} else memcpy(st, input + c * n_in + n_in - hist, sizeof(float) * hist);
	}
	return SBM_OK;
}

// CreateShHrirsFromWav's resampling step (RA/dsp/sh_hrir_creator.cc:42-55): SH-HRIRs [channels][taps] at source_hz -> the same at
// destination_hz, *out_taps = Resampler::GetNextOutputLength(taps) long.  `out` holds out_cap_taps taps per channel.
SBM_API int sbm_resample_sh_hrir(int device, const float* sh_hrir, int channels, int taps, int source_hz, int destination_hz, float* out,
                                 int out_cap_taps, int* out_taps) {
	if (!sh_hrir || channels <= 0 || taps <= 0 || !out_taps) return SBM_ERR_INVALID_ARG;
	if (!sbm_polyphase_rates_supported(source_hz, destination_hz)) return SBM_ERR_UNSUPPORTED;
	sbm_polyphase* p = nullptr;
	int rc = sbm_polyphase_create(device, &p);
	if (rc) return rc;
	sbm_polyphase_reset_state(p);
	rc = sbm_polyphase_set_rate_and_channels(p, source_hz, destination_hz, channels);
	const size_t n_out = next_output_length(p, (size_t)taps);
	*out_taps = (int)n_out;
	if (!rc && (!out || (size_t)out_cap_taps < n_out)) rc = out ? SBM_ERR_CAPACITY : SBM_OK;      // out == NULL: length query
	else if (!rc) {
		std::vector<float> tmp((size_t)channels * n_out);
		rc = sbm_polyphase_process(p, sh_hrir, (size_t)taps, tmp.data(), n_out);
		if (!rc) for (int c = 0; c < channels; ++c) memcpy(out + (size_t)c * (size_t)out_cap_taps, tmp.data() + (size_t)c * n_out, sizeof(float) * n_out);
	}
	sbm_polyphase_destroy(p);
	return rc;
}

}  // extern "C"
