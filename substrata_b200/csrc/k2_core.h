// k2_core.h -- per-sample arithmetic of the mix kernel ("K2") and of the finish kernel ("K4"), as
// __host__ __device__ functions so that the CUDA kernels (k2_mix.cu, device_cuda.cu) and the host emulation used
// by the no-GPU test suite (tests/native/device_emu.cc -- test infrastructure, never shipped) evaluate literally the
// same expressions.  The kernels add only the parallel decomposition (who owns which frame, how partial buses are
// reduced); the optimised kernel paths in k2_mix.cu are checked against these functions by the GPU parity tests.
//
// Follows: the feeder's gather (AE.cpp:389-432), ring pop (AE.cpp:459-474), pitch/doppler mixer (AE.cpp:435-458),
// AudioResampler::resampleFromContiguousBuffer (AR.cpp:57-74), LinearGainRamp / ConstantGain (RA/dsp/gain.cc:25-92),
// MonoPoleFilter::Filter (RA/dsp/mono_pole_filter.cc:34-57), AmbisonicBinauralDecoder::Process
// (RA/ambisonics/ambisonic_binaural_decoder.cc:45-78).
#pragma once
#include "sbm_core.h"

namespace sbm {

#if defined(__CUDA_ARCH__)
SBM_HD float ld_pcm(const float* p) { return __ldg(p); }
#else
SBM_HD float ld_pcm(const float* p) { return *p; }
#endif

// Sample j of the window the feeder gathered for this buffer: pcm[start + j] with loop / ring wrap, zeros past
// `avail` (end of a non-looping buffer, ring underflow).
SBM_HD float window_sample(const WorkRec& r, uint32_t j) {
	if (j >= r.avail) return 0.0f;
	uint32_t idx = r.start + j;
	if (r.flags & kRecWrap) {
		if (idx >= r.len) idx -= r.len;
		if (idx >= r.len) idx %= r.len;     // loops shorter than one buffer
	}
	return ld_pcm(r.pcm + idx);
}

// Element j of AudioResampler's temp buffer {prev_samples[0], prev_samples[1], window...} (AR.cpp:95-99).
SBM_HD float resampler_tmp(const WorkRec& r, int j) {
	if (j == 0) return r.prev0;
	if (j == 1) return r.prev1;
	return window_sample(r, (uint32_t)(j - 2));
}

// Read-head coordinate of output sample i in 32.32 fixed point -> integer sample index d and interpolation fraction.
// The fraction keeps the top 23 bits (truncation error < 1.2e-7 of a sample; the interpolant is continuous in it).
SBM_HD void fx_split(unsigned long long x0_fx, unsigned long long step_fx, uint32_t i, int* d, float* frac) {
	const unsigned long long p = x0_fx + (unsigned long long)i * step_fx;
	*d = (int)(uint32_t)(p >> 32);
	*frac = fsub(u2f(0x3f800000u | ((uint32_t)p >> 9)), 1.0f);
}

// Output sample i of the linear sample-rate conversion (AR.cpp:57-74): out = tmp[d] (1 - frac) + tmp[d + 1] frac with
// clamped indices.  The reference advances the coordinate by repeated double addition; the fixed-point head differs
// from it by < 1e-6 of a sample over a buffer.
SBM_HD float resampled_sample(const WorkRec& r, uint32_t i) {
	int d; float frac;
	fx_split(r.x0_fx, r.step_fx, i, &d, &frac);
	const int max_i = (int)r.n_src + 1;
	const int d0 = d > max_i ? max_i : d;
	const int d1 = d + 1 > max_i ? max_i : d + 1;
	return fadd(fmul(resampler_tmp(r, d0), fsub(1.0f, frac)), fmul(resampler_tmp(r, d1), frac));
}

// Output sample i of a single pitch-shifted mix source in window form (AE.cpp:443-452; K1 turned the read head into
// window + fixed-point coordinate): sample = w[d] (1 - frac) + w[d + 1] frac, times mix_factor.
SBM_HD float mix1_sample(const WorkRec& r, uint32_t i) {
	int d; float frac;
	fx_split(r.x0_fx, r.step_fx, i, &d, &frac);
	const float s = fadd(fmul(window_sample(r, (uint32_t)d), fsub(1.0f, frac)), fmul(window_sample(r, (uint32_t)d + 1u), frac));
	return fmul(s, r.mix_factor);
}

// Output sample i of the general pitch / doppler mixer (AE.cpp:435-458), any number of mix sources: the read head of
// every mix source advances by source_delta BEFORE each read.
SBM_HD float mixed_sample(const WorkRec& r, const MixZ* mixz, const double* mix_pos0, uint32_t i) {
	float acc = 0.0f;
	for (uint32_t z = 0; z < r.nz; ++z) {
		const MixZ m = mixz[r.mix_first + z];
		const double pos = fma((double)(i + 1), m.delta, mix_pos0[r.mix_first + z]);
		const unsigned long long ip = (unsigned long long)pos;
		const uint32_t idx = (uint32_t)(ip % m.len);
		const uint32_t idx1 = idx + 1 == m.len ? 0 : idx + 1;
		const float frac = (float)(pos - (double)ip);
		const float s = fadd(fmul(ld_pcm(m.pcm + idx), fsub(1.0f, frac)), fmul(ld_pcm(m.pcm + idx1), frac));
		acc = fadd(acc, fmul(s, m.mix_factor));
	}
	return acc;
}

// Engine-rate input sample i of a source (what SetPlanarBuffer receives, AE.cpp:480-494).
SBM_HD float source_sample(const WorkRec& r, const MixZ* mixz, const double* mix_pos0, uint32_t i) {
	if (r.flags & kRecMix1) return mix1_sample(r, i);
	if (r.flags & kRecMix) return mixed_sample(r, mixz, mix_pos0, i);
	if (r.flags & kRecResample) return resampled_sample(r, i);
	return window_sample(r, i);
}

// Gain at sample i of a planned ramp: the reference's running `gain += inc` (RA/dsp/gain.cc:38-47) evaluated in
// closed form inside the buffer (<= 0.5 ulp * sqrt(i) away from the running sum), the exact end value afterwards.
SBM_HD float gain_at(const RampDesc& d, uint32_t i) {
	return i < d.n ? fmaf((float)i, d.inc, d.start) : d.konst;
}

// Gain of ambisonic bus channel c at frame i (ShRamps, sbm_core.h): constant part + the ramp's offset from it while it lasts.
SBM_HD float sh_gain_at(float konst, const ShRamps* rp, int c, uint32_t i) {
	if (!rp) return konst;
	const float fi = (float)i;
	return fi < rp->nf[c] ? fadd(konst, fmaf(fi, rp->inc[c], rp->stm[c])) : konst;
}

// New AudioResampler::prev_samples after this buffer (AR.cpp:108-110): the last two of {prev0, prev1, window}.
SBM_HD void resampler_tail(const WorkRec& r, float* p0, float* p1) {
	const int n_tmp = (int)r.n_src + 2;
	*p0 = resampler_tmp(r, n_tmp - 2);
	*p1 = resampler_tmp(r, n_tmp - 1);
}

// One step of the occlusion low-pass (RA/dsp/mono_pole_filter.cc:47-53).
SBM_HD float lpf_step(float coef, float prev_out, float in) { return fadd(fmul(coef, fsub(prev_out, in)), in); }

// ACN channel c has negative degree m  <=>  the right ear subtracts it (RA/ambisonics/ambisonic_binaural_decoder.cc:60-76,
// RA/ambisonics/utils.h:102-105).
SBM_HD bool sh_right_ear_negated(int c) {
	int n = 0;
	while ((n + 1) * (n + 1) <= c) ++n;
	return c - n * (n + 1) < 0;
}

// Sample `j` (may be negative: history) of bus channel c as the decoder's FIR sees it.
SBM_HD float decode_input(const float* bus_c, const float* hist_c, int H, int j) {
	if (j >= 0) return bus_c[j];
	return H + j >= 0 ? hist_c[H + j] : 0.0f;
}

// SH-domain binaural decode of frame i: causal linear convolution of every bus channel with its SH-HRIR (direct
// form, double accumulation -- the reference's partitioned FFT computes the same convolution in f32), summed into L
// and, sign by ACN degree, into R.
SBM_HD void decode_frame(const float* bus, const float* hrir, const float* hist, uint32_t frames, uint32_t taps, uint32_t i,
                         float* out_l, float* out_r, int nch = kNumSh) {
	const int H = (int)taps - 1;
	float l = 0.0f, r = 0.0f;
	for (int c = 0; c < nch; ++c) {
		const float* h = hrir + (size_t)c * taps;
		const float* x = bus + (size_t)c * frames;
		const float* hc = hist + (size_t)c * (size_t)(H > 0 ? H : 1);
		double acc = 0.0;
		for (int t = 0; t < (int)taps; ++t) acc = fma((double)h[t], (double)decode_input(x, hc, H, (int)i - t), acc);
		const float y = (float)acc;
		l = fadd(l, y);
		r = sh_right_ear_negated(c) ? fsub(r, y) : fadd(r, y);
	}
	*out_l = l; *out_r = r;
}

// What the finish step does this buffer, decided by one thread from the (possibly cross-rank summed) bus flags
// {ambisonic mixer used, stereo panner used, stereo-source mixer used}:
//   AmbisonicBinauralDecoderNode::AudioProcess (RA/graph/ambisonic_binaural_decoder_node.cc:56-109): decode when the
//   bus has input, else flush the FIR tail until `frames` silent frames have been pushed through;
//   FillInterleavedOutputBuffer returns false when nothing reached the stereo mixer (resonance_audio_api_impl.cc:498-522).
// Also publishes the global "mixer was used" bits for the next buffer's GainMixer sweep (k1_core.h) and clears the
// other parity for K1's atomicOr.
//   Room effects: the order-1 decoder of the reflections follows the same decode / flush rule on the ReflectionsNode's output
//   (DevFlags::foa_valid, set by the room kernel); the ReverbNode -- gain 0 with Substrata's settings -- contributes an all-zero
//   buffer whenever the reverb send mixer was fed, which only shows in `valid` (RA/graph/reverb_node.cc:118-145).
struct FinishPlan { int sh, stereo, decode, decode1, valid; };
SBM_HD bool room_decode1_wanted(const DevFlags* f, uint32_t frames) { return f->foa_valid || f->sh1_flush_frames < (int)frames; }
SBM_HD FinishPlan plan_finish(const float* bus_flags, bool have_hrir, bool have_hrir1, bool room_enabled, DevFlags* f, uint32_t parity, uint32_t frames) {
	FinishPlan p;
	const int sh = bus_flags[0] > 0.0f, pan = bus_flags[1] > 0.0f, st = bus_flags[2] > 0.0f, refl = bus_flags[3] > 0.0f;
	p.sh = sh; p.stereo = pan || st; p.decode = 0; p.decode1 = 0;
	if (have_hrir) {
		if (sh) { f->sh_flush_frames = 0; p.decode = 1; }
		else if (f->sh_flush_frames < (int)frames) { f->sh_flush_frames += (int)frames; p.decode = 1; }
	}
	if (have_hrir1) {
		if (f->foa_valid) { f->sh1_flush_frames = 0; p.decode1 = 1; }
		else if (f->sh1_flush_frames < (int)frames) { f->sh1_flush_frames += (int)frames; p.decode1 = 1; }
	}
	f->foa_valid = 0;
	p.valid = p.stereo || p.decode || p.decode1 || (room_enabled && refl);
	f->sh_used[parity] = sh; f->pan_used[parity] = pan; f->st_used[parity] = st; f->refl_used[parity] = refl;
	f->sh_used[parity ^ 1] = 0; f->pan_used[parity ^ 1] = 0; f->st_used[parity ^ 1] = 0; f->refl_used[parity ^ 1] = 0;
	return p;
}

}  // namespace sbm
