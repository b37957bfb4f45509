// oracle/ref_driver.cc -- TEST INFRASTRUCTURE ONLY.  Never linked into libsubstrata_b200.so.
//
// Drives the REAL reference implementation of the mix path, compiled in place from /root/reference
// by oracle/Makefile into oracle/_ref/libsubstrata_ref.so:
//   * vraudio::ResonanceAudioApi            (resonance-audio/resonance_audio/api/resonance_audio_api.h)
//   * glare::AudioResampler                 (audio/AudioResampler.cpp, unmodified, shim headers)
// audio/AudioEngine.cpp itself cannot be compiled here (glare-core is not vendored and init() opens
// an audio device), so the per-buffer feeder body of ResonanceThread::doRun (AudioEngine.cpp:367-537)
// is restated below around the real Resonance object.  The C entry points mirror include/sbm_audio.h
// one to one so the parity tests can drive both with the same scenario.
//
// Used by: tests/ (as the checker) and bench.py --impl reference / cpu_baseline (as the timed CPU arm).

#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "api/resonance_audio_api.h"
#include "ambisonics/ambisonic_binaural_decoder.h"
#include "ambisonics/ambisonic_lookup_table.h"
#include "base/audio_buffer.h"
#include "base/spherical_angle.h"
#include "dsp/fft_manager.h"
#include "dsp/resampler.h"
#include "dsp/sh_hrir_creator.h"
#include "graph/resonance_audio_api_impl.h"

#include "AudioResampler.h"  // the reference's own class (audio/AudioResampler.h)

namespace {

struct RefBuffer { std::vector<float> samples; };

struct RefMixSource {           // glare::MixSource, AudioEngine.h:61-69
	int buf;
	double sound_file_i;
	double source_delta;
	float mix_factor;
};

struct RefSource {              // the feeder-relevant part of glare::AudioSource, AudioEngine.h:83-155
	int handle;                 // Resonance source id
	int kind;
	int type;                   // 0 = NonStreaming (shared_buffer), 1 = Streaming ring, 2 = Streaming mix_sources
	int sampling_rate;
	int buf;                    // shared_buffer
	size_t cur_read_i;
	bool looping, remove_on_finish;
	int64_t eof_marker;         // EOF_marker_position
	std::deque<float> ring;     // CircularBuffer<float> buffer
	std::vector<RefMixSource> mix;
	glare::AudioResampler resampler;
	bool finished;
};

struct RefEngine {
	vraudio::ResonanceAudioApi* api;
	int frames, rate;
	std::vector<RefBuffer> buffers;
	std::map<int, RefSource> sources;   // ordered by id; AudioEngine iterates a std::set of pointers (arbitrary order)
	std::vector<int> finished_ids;
	js::Vector<float, 16> temp_buf, temp_resampling_buf, resampled_buf;
	// Optional truncated-HRIR decode: the reference's own AmbisonicBinauralDecoder fed
	// with the first `hrir_taps` taps of the stock SH-HRIRs, run on the real 16-channel bus.
	int hrir_taps;
	std::unique_ptr<vraudio::FftManager> fft;
	std::unique_ptr<vraudio::AmbisonicBinauralDecoder> trunc_decoder;
	std::unique_ptr<vraudio::AudioBuffer> trunc_out;
	std::vector<float> last_bus;        // 16 x frames, valid after mix when has_bus
	bool has_bus;
	int trunc_flush_frames;
};

const int kTypeBuffer = 0, kTypeRing = 1, kTypeMix = 2;

}  // namespace

extern "C" {

void* ref_create(int frames_per_buffer, int sample_rate_hz, int hrir_taps) {
	RefEngine* e = new RefEngine();
	e->api = vraudio::CreateResonanceAudioApi(2, (size_t)frames_per_buffer, sample_rate_hz);  // AudioEngine.cpp:676
	e->frames = frames_per_buffer;
	e->rate = sample_rate_hz;
	e->hrir_taps = hrir_taps;
	e->has_bus = false;
	e->trunc_flush_frames = frames_per_buffer;
	// Substrata disables room effects outside parcels (gui_client/GUIClient.cpp:6896); the benchmark
	// configurations run in that state.
	e->api->EnableRoomEffects(false);
	if (hrir_taps > 0) {
		vraudio::Resampler resampler;
		std::unique_ptr<vraudio::AudioBuffer> full =
			vraudio::CreateShHrirsFromAssets("WAV/Subject_002/SH/sh_hrir_order_3.wav", sample_rate_hz, &resampler);
		const size_t taps = std::min((size_t)hrir_taps, full->num_frames());
		vraudio::AudioBuffer cut(full->num_channels(), taps);
		for (size_t c = 0; c < full->num_channels(); ++c)
			for (size_t t = 0; t < taps; ++t) cut[c][t] = (*full)[c][t];
		e->fft.reset(new vraudio::FftManager((size_t)frames_per_buffer));
		e->trunc_decoder.reset(new vraudio::AmbisonicBinauralDecoder(cut, (size_t)frames_per_buffer, e->fft.get()));
		e->trunc_out.reset(new vraudio::AudioBuffer(2, (size_t)frames_per_buffer));
	}
	return e;
}

void ref_destroy(void* h) {
	RefEngine* e = (RefEngine*)h;
	delete e->api;   // AudioEngine.cpp:837
	delete e;
}

int ref_buffer_create(void* h, const float* samples, size_t n) {
	RefEngine* e = (RefEngine*)h;
	e->buffers.emplace_back();
	e->buffers.back().samples.assign(samples, samples + n);
	return (int)e->buffers.size() - 1;
}

// kind: 0 = kStereoPanning sound object, 3 = kBinauralHighQuality sound object (what AudioEngine.cpp:857
// hard-codes), 100 = non-spatial stereo source (AudioEngine.cpp:867).
int ref_source_create(void* h, int kind) {
	RefEngine* e = (RefEngine*)h;
	int handle;
	if (kind == 100) handle = e->api->CreateStereoSource(2);
	else handle = e->api->CreateSoundObjectSource((vraudio::RenderingMode)kind);
	RefSource& s = e->sources[handle];
	s.handle = handle; s.kind = kind; s.type = kTypeBuffer; s.sampling_rate = e->rate; s.buf = -1; s.cur_read_i = 0;
	s.looping = true; s.remove_on_finish = false; s.eof_marker = std::numeric_limits<int64_t>::max(); s.finished = false;
	s.resampler.init(e->rate, e->rate);
	// Creating a source posts a task on Resonance's bounded queue (25 600 entries, silently dropping:
	// resonance_audio_api_impl.cc:41, lockless_task_queue.cc:43-51).  Drain it periodically the way a
	// running ResonanceThread would (SetPlanarBuffer executes the queue, :538).
	if ((handle & 1023) == 1023) { const float* p = nullptr; e->api->SetPlanarBuffer(-1, &p, 1, (size_t)e->frames); }
	return handle;
}

static void drain(RefEngine* e) { const float* p = nullptr; e->api->SetPlanarBuffer(-1, &p, 1, (size_t)e->frames); }

void ref_source_attach_buffer(void* h, int id, int buf, uint64_t cur_read_i, int sampling_rate, int looping, int remove_on_finish) {
	RefEngine* e = (RefEngine*)h; RefSource& s = e->sources.at(id);
	s.type = kTypeBuffer; s.buf = buf; s.cur_read_i = (size_t)cur_read_i; s.sampling_rate = sampling_rate;
	s.looping = looping != 0; s.remove_on_finish = remove_on_finish != 0;
	s.resampler.init(sampling_rate, e->rate);   // AudioEngine.cpp:875
}

void ref_source_attach_ring(void* h, int id, int sampling_rate) {
	RefEngine* e = (RefEngine*)h; RefSource& s = e->sources.at(id);
	s.type = kTypeRing; s.sampling_rate = sampling_rate; s.resampler.init(sampling_rate, e->rate);
}

void ref_source_ring_push(void* h, int id, const float* samples, size_t n) {
	RefEngine* e = (RefEngine*)h; RefSource& s = e->sources.at(id);
	s.ring.insert(s.ring.end(), samples, samples + n);
}

void ref_source_attach_mix(void* h, int id, int sampling_rate, int nz, const int* bufs, const double* sound_file_i,
                           const double* source_delta, const float* mix_factor) {
	RefEngine* e = (RefEngine*)h; RefSource& s = e->sources.at(id);
	s.type = kTypeMix; s.sampling_rate = sampling_rate; s.resampler.init(sampling_rate, e->rate);
	s.mix.resize((size_t)nz);
	for (int z = 0; z < nz; ++z) s.mix[(size_t)z] = RefMixSource{bufs[z], sound_file_i[z], source_delta[z], mix_factor[z]};
}

void ref_source_set_mix_params(void* h, int id, int z, double source_delta, float mix_factor) {
	RefEngine* e = (RefEngine*)h; RefSource& s = e->sources.at(id);
	s.mix.at((size_t)z).source_delta = source_delta; s.mix.at((size_t)z).mix_factor = mix_factor;
}

void ref_source_set_eof_marker(void* h, int id, int64_t samples_to_eof) { ((RefEngine*)h)->sources.at(id).eof_marker = samples_to_eof; }
void ref_source_set_looping(void* h, int id, int looping) { ((RefEngine*)h)->sources.at(id).looping = looping != 0; }

static int g_setter_count = 0;
static inline void maybe_drain(RefEngine* e) { if (++g_setter_count >= 8192) { g_setter_count = 0; drain(e); } }

void ref_set_source_positions(void* h, int n, const int* ids, const float* xyz) {
	RefEngine* e = (RefEngine*)h;
	for (int i = 0; i < n; ++i) { e->api->SetSourcePosition(ids[i], xyz[3*i], xyz[3*i+1], xyz[3*i+2]); maybe_drain(e); }  // AudioEngine.cpp:955
}
void ref_set_source_volumes(void* h, int n, const int* ids, const float* vol) {
	RefEngine* e = (RefEngine*)h;
	for (int i = 0; i < n; ++i) { e->api->SetSourceVolume(ids[i], vol[i]); maybe_drain(e); }                                // AudioEngine.cpp:965
}
void ref_set_source_occlusions(void* h, int n, const int* ids, const float* occ) {
	RefEngine* e = (RefEngine*)h;
	for (int i = 0; i < n; ++i) { e->api->SetSoundObjectOcclusionIntensity(ids[i], occ[i]); maybe_drain(e); }              // AudioEngine.cpp:974
}
void ref_set_head(void* h, const float* pos, const float* quat_xyzw) {
	RefEngine* e = (RefEngine*)h;
	e->api->SetHeadPosition(pos[0], pos[1], pos[2]);                                   // AudioEngine.cpp:987
	e->api->SetHeadRotation(quat_xyzw[0], quat_xyzw[1], quat_xyzw[2], quat_xyzw[3]);   // AudioEngine.cpp:988
}
void ref_set_master_volume(void* h, float v) { ((RefEngine*)h)->api->SetMasterVolume(v); }  // AudioEngine.cpp:810

void ref_source_destroy(void* h, int id) {
	RefEngine* e = (RefEngine*)h;
	auto it = e->sources.find(id);
	if (it == e->sources.end()) return;
	e->api->DestroySource(id);   // AudioEngine.cpp:908
	e->sources.erase(it);
}

// One pass of the body of `while(num_samples_buffered < ...)` in ResonanceThread::doRun
// (AudioEngine.cpp:365-537) without the first-second mute (:532).  Returns what
// FillInterleavedOutputBuffer returned.
int ref_mix(void* h, float* out_interleaved) {
	RefEngine* e = (RefEngine*)h;
	const size_t F = (size_t)e->frames;
	std::vector<int> to_erase;
	for (auto& kv : e->sources) {
		RefSource& s = kv.second;
		bool remove_source = false;
		size_t n_src = F;                                                        // AudioEngine.cpp:378-382
		if (s.sampling_rate != e->rate) n_src = (size_t)(int)s.resampler.numSrcSamplesNeeded(F);
		e->temp_buf.resizeNoCopy(n_src);
		const float* contiguous = e->temp_buf.data();
		if (s.type == kTypeBuffer) {                                              // AudioEngine.cpp:389-432
			const std::vector<float>& sb = e->buffers.at((size_t)s.buf).samples;
			if (s.cur_read_i + n_src <= sb.size()) {
				contiguous = sb.data() + s.cur_read_i;
				s.cur_read_i += n_src;
				if (s.looping && s.cur_read_i == sb.size()) s.cur_read_i = 0;
			} else {
				size_t cur = s.cur_read_i;
				for (size_t i = 0; i < n_src; ++i) {
					if (cur < sb.size()) {
						e->temp_buf[i] = sb[cur++];
						if (s.looping && cur == sb.size()) cur = 0;
					} else e->temp_buf[i] = 0.f;
				}
				s.cur_read_i = cur;
				remove_source = s.remove_on_finish && (cur >= sb.size());
			}
		} else if (s.type == kTypeMix) {                                          // AudioEngine.cpp:435-458
			for (size_t i = 0; i < n_src; ++i) e->temp_buf[i] = 0.f;
			for (RefMixSource& m : s.mix) {
				const std::vector<float>& sb = e->buffers.at((size_t)m.buf).samples;
				const size_t L = sb.size();
				for (size_t i = 0; i < n_src; ++i) {
					m.sound_file_i += m.source_delta;
					const size_t index = (size_t)m.sound_file_i % L;
					const size_t index_1 = (index + 1) % L;
					const float frac = (float)(m.sound_file_i - (size_t)m.sound_file_i);
					const float sample = sb[index] * (1 - frac) + sb[index_1] * frac;
					e->temp_buf[i] += sample * m.mix_factor;
				}
			}
		} else {                                                                  // AudioEngine.cpp:459-474
			const size_t have = std::min(s.ring.size(), n_src);
			for (size_t i = 0; i < have; ++i) e->temp_buf[i] = s.ring[i];
			s.ring.erase(s.ring.begin(), s.ring.begin() + (long)have);
			for (size_t i = have; i < n_src; ++i) e->temp_buf[i] = 0.f;
		}
		if (s.sampling_rate == e->rate) {                                         // AudioEngine.cpp:480-494
			e->api->SetPlanarBuffer(s.handle, &contiguous, 1, F);
		} else {
			e->resampled_buf.resizeNoCopy(F);
			s.resampler.resample(e->resampled_buf.data(), F, contiguous, n_src, e->temp_resampling_buf);
			const float* p = e->resampled_buf.data();
			e->api->SetPlanarBuffer(s.handle, &p, 1, F);
		}
		s.eof_marker -= (int64_t)n_src;                                           // AudioEngine.cpp:496-517
		const bool reached_eof = s.eof_marker <= 0;
		if (remove_source || (!s.looping && reached_eof)) {
			e->api->DestroySource(s.handle);
			e->finished_ids.push_back(s.handle);
			to_erase.push_back(s.handle);
		}
	}
	for (int id : to_erase) e->sources.erase(id);

	const bool ok = e->api->FillInterleavedOutputBuffer(2, F, out_interleaved);  // AudioEngine.cpp:522-526
	if (!ok) for (size_t i = 0; i < 2 * F; ++i) out_interleaved[i] = 0.f;

	// 16-channel ambisonic bus of this buffer (order-3 mixer node output), for stage-wise bisecting and the
	// truncated-tap decode.
	const vraudio::AudioBuffer* bus = static_cast<vraudio::ResonanceAudioApiImpl*>(e->api)->GetAmbisonicOutputBuffer();
	e->has_bus = bus != nullptr;
	if (bus) {
		e->last_bus.assign(16 * F, 0.f);
		for (size_t c = 0; c < std::min<size_t>(16, bus->num_channels()); ++c)
			for (size_t i = 0; i < F; ++i) e->last_bus[c * F + i] = (*bus)[c][i];
	}
	if (e->trunc_decoder) {
		// Same node-level behaviour as AmbisonicBinauralDecoderNode (ambisonic_binaural_decoder_node.cc:56-80):
		// decode when there is input, flush the tail once with silence, then report "no output".
		vraudio::AudioBuffer in(16, F);
		in.Clear();
		bool run = false;
		if (bus) { for (size_t c = 0; c < 16; ++c) for (size_t i = 0; i < F; ++i) in[c][i] = (*bus)[c][i]; e->trunc_flush_frames = 0; run = true; }
		else if (e->trunc_flush_frames < (int)F) { e->trunc_flush_frames += (int)F; run = true; }
		if (run) {
			e->trunc_decoder->Process(in, e->trunc_out.get());
			for (size_t i = 0; i < F; ++i) { out_interleaved[2*i] = (*e->trunc_out)[0][i]; out_interleaved[2*i+1] = (*e->trunc_out)[1][i]; }
			return 1;
		}
		for (size_t i = 0; i < 2 * F; ++i) out_interleaved[i] = 0.f;
		return 0;
	}
	return ok ? 1 : 0;
}

int ref_get_ambisonic_bus(void* h, float* out16xF) {
	RefEngine* e = (RefEngine*)h;
	if (!e->has_bus) return 0;
	std::memcpy(out16xF, e->last_bus.data(), e->last_bus.size() * sizeof(float));
	return 1;
}

int ref_poll_finished(void* h, int* ids, int cap) {
	RefEngine* e = (RefEngine*)h;
	int n = 0;
	for (; n < cap && n < (int)e->finished_ids.size(); ++n) ids[n] = e->finished_ids[(size_t)n];
	e->finished_ids.erase(e->finished_ids.begin(), e->finished_ids.begin() + n);
	return n;
}

// cur_read_i of a buffer source / remaining ring samples / mix read head, for feeder-state parity (row a17).
int ref_source_query(void* h, int id, uint64_t* cur_read_i, int64_t* eof_marker, uint64_t* ring_size, double* mix_pos0) {
	RefEngine* e = (RefEngine*)h;
	auto it = e->sources.find(id);
	if (it == e->sources.end()) return 0;
	*cur_read_i = it->second.cur_read_i; *eof_marker = it->second.eof_marker; *ring_size = it->second.ring.size();
	*mix_pos0 = it->second.mix.empty() ? 0.0 : it->second.mix[0].sound_file_i;
	return 1;
}

// Room effects as Substrata drives them: setRoomEffectsEnabled -> EnableRoomEffects (AudioEngine.cpp:738-743),
// setCurentRoomDimensions -> SetReflectionProperties (AudioEngine.cpp:746-768).
void ref_enable_room_effects(void* h, int enable) { ((RefEngine*)h)->api->EnableRoomEffects(enable != 0); }
void ref_set_reflection_properties(void* h, const float* room_position, const float* room_rotation_xyzw, const float* room_dimensions,
                                   const float* coefficients6, float cutoff_frequency, float gain) {
	vraudio::ReflectionProperties p;
	for (int i = 0; i < 3; ++i) { p.room_position[i] = room_position[i]; p.room_dimensions[i] = room_dimensions[i]; }
	for (int i = 0; i < 4; ++i) p.room_rotation[i] = room_rotation_xyzw[i];
	for (int i = 0; i < 6; ++i) p.coefficients[i] = coefficients6[i];
	p.cutoff_frequency = cutoff_frequency; p.gain = gain;
	((RefEngine*)h)->api->SetReflectionProperties(p);
}

// ---- stage-level probes of the real reference, used to pin the oracle restatement -----------------------
// The stock SH-HRIRs of ambisonic order 1 (the reflections' decoder) exactly as that decoder receives them.
int ref_dump_sh_hrirs_order1(int sample_rate_hz, float* out, int cap_floats, int* channels) {
	vraudio::Resampler resampler;
	std::unique_ptr<vraudio::AudioBuffer> hr =
		vraudio::CreateShHrirsFromAssets("WAV/Subject_002/SH/sh_hrir_order_1.wav", sample_rate_hz, &resampler);
	const int C = (int)hr->num_channels(), T = (int)hr->num_frames();
	*channels = C;
	if (out && cap_floats >= C * T)
		for (int c = 0; c < C; ++c) for (int t = 0; t < T; ++t) out[c * T + t] = (*hr)[(size_t)c][(size_t)t];
	return T;
}

// The stock order-3 SH-HRIRs exactly as the decoder receives them (sh_hrir_creator.cc:60-70).  Returns taps.
int ref_dump_sh_hrirs(int sample_rate_hz, float* out, int cap_floats, int* channels) {
	vraudio::Resampler resampler;
	std::unique_ptr<vraudio::AudioBuffer> hr =
		vraudio::CreateShHrirsFromAssets("WAV/Subject_002/SH/sh_hrir_order_3.wav", sample_rate_hz, &resampler);
	const int C = (int)hr->num_channels(), T = (int)hr->num_frames();
	*channels = C;
	if (out && cap_floats >= C * T)
		for (int c = 0; c < C; ++c) for (int t = 0; t < T; ++t) out[c * T + t] = (*hr)[(size_t)c][(size_t)t];
	return T;
}

// GetEncodingCoeffs at an arbitrary direction (ambisonic_lookup_table.cc:88-161), order 3, given spread.
void ref_encoding_coeffs(float azimuth_rad, float elevation_rad, float spread_deg, float* out16) {
	static vraudio::AmbisonicLookupTable table(3);
	std::vector<float> c(16);
	table.GetEncodingCoeffs(3, vraudio::SphericalAngle(azimuth_rad, elevation_rad), spread_deg, &c);
	for (int i = 0; i < 16; ++i) out16[i] = c[(size_t)i];
}

// glare::AudioResampler driven stand-alone (the reference's AudioResampler::test() pattern, AudioResampler.cpp:213-287).
void* ref_resampler_create(int src_rate, int dst_rate) { auto* r = new glare::AudioResampler(); r->init(src_rate, dst_rate); return r; }
void ref_resampler_destroy(void* r) { delete (glare::AudioResampler*)r; }
size_t ref_resampler_num_src_needed(void* r, size_t n_dst) { return ((glare::AudioResampler*)r)->numSrcSamplesNeeded(n_dst); }
void ref_resampler_resample(void* r, float* dst, size_t n_dst, const float* src, size_t n_src) {
	js::Vector<float, 16> tmp;
	((glare::AudioResampler*)r)->resample(dst, n_dst, src, n_src, tmp);
}

float ref_atan2f(float y, float x) { return std::atan2(y, x); }   // the libm the reference is linked against

// vraudio::Resampler (the polyphase resampler, RA/dsp/resampler.h) driven stand-alone: planar [channels][frames] in and out.
void* ref_polyphase_create() { return new vraudio::Resampler(); }
void ref_polyphase_destroy(void* p) { delete (vraudio::Resampler*)p; }
int ref_polyphase_rates_supported(int s, int d) { return vraudio::Resampler::AreSampleRatesSupported(s, d) ? 1 : 0; }
void ref_polyphase_set_rate_and_channels(void* p, int s, int d, int nch) { ((vraudio::Resampler*)p)->SetRateAndNumChannels(s, d, (size_t)nch); }
void ref_polyphase_reset_state(void* p) { ((vraudio::Resampler*)p)->ResetState(); }
size_t ref_polyphase_max_output_length(void* p, size_t n) { return ((vraudio::Resampler*)p)->GetMaxOutputLength(n); }
size_t ref_polyphase_next_output_length(void* p, size_t n) { return ((vraudio::Resampler*)p)->GetNextOutputLength(n); }
void ref_polyphase_process(void* p, const float* input, size_t n_in, int nch, float* output, size_t n_out) {
	vraudio::AudioBuffer in((size_t)nch, n_in), out((size_t)nch, n_out);
	for (int c = 0; c < nch; ++c) for (size_t i = 0; i < n_in; ++i) in[(size_t)c][i] = input[(size_t)c * n_in + i];
	((vraudio::Resampler*)p)->Process(in, &out);
	for (int c = 0; c < nch; ++c) for (size_t i = 0; i < n_out; ++i) output[(size_t)c * n_out + i] = out[(size_t)c][i];
}

}  // extern "C"
