// tests/native/test_streaming_host.cc -- TEST INFRASTRUCTURE.  The glare::AudioEngine mirror's streaming boundary
// (addSourceFromStreamingSoundFile, needNewStreamerForPath, the stream maps, sync-start of sources sharing a stream,
// StreamerThread::streamOnce, EOF markers, a decoder that reports its own sampling rate, seekToStartAndUnpauseAudio) on the
// emulated device, against the oracle fed with exactly the samples the synthetic decoder produced.
// Reference behaviour: audio/AudioEngine.cpp:1081-1146,770-803,898-943; audio/StreamerThread.cpp:27-120.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <limits>
#include <vector>

#include "AudioEngine.h"
#include "StreamerThread.h"
extern "C" {
#include "../../oracle/mix_oracle.h"
}

// the batched AudioResampler drop-in lives in the CUDA library only; this binary never resamples outside the engine
extern "C" {
int sbm_resampler_create(int, int, const int*, int, sbm_resampler**) { return SBM_ERR_CUDA; }
void sbm_resampler_destroy(sbm_resampler*) {}
int sbm_resampler_num_src_needed(sbm_resampler*, size_t, size_t*) { return SBM_ERR_CUDA; }
int sbm_resampler_resample(sbm_resampler*, float*, size_t, const float*, const size_t*, size_t) { return SBM_ERR_CUDA; }
}

using namespace glare;

// A decoder that produces 1152-sample-frame stereo frames of a deterministic signal at `rate` Hz, `total_frames` frames per pass.
class SynthStreamer : public MP3AudioStreamer {
public:
	SynthStreamer(int rate_, int total_frames_, float f0_) : rate(rate_), total_frames(total_frames_), frame(0), f0(f0_), seeks(0) {}
	bool decodeFrame(js::Vector<float, 16>& out, int& nch, int& hz) override {
		out.resizeNoCopy(1152 * 2);
		for (int i = 0; i < 1152; ++i) {
			const double t = (double)((size_t)frame * 1152 + i) / rate;
			const float l = 0.4f * (float)sin(2 * M_PI * f0 * t), r = 0.2f * (float)sin(2 * M_PI * 1.5 * f0 * t + 0.3);
			out[2 * i] = l; out[2 * i + 1] = r;
		}
		nch = 2; hz = rate;
		++frame;
		return frame >= total_frames;
	}
	void seekToBeginningOfFile() override { frame = 0; ++seeks; }
	void seekToApproxTimeWrapped(double global_time) override { frame = (int)(global_time * rate / 1152.0) % total_frames; }
	int rate, total_frames, frame; float f0; int seeks;
};

static double rel_rms(const std::vector<float>& a, const std::vector<float>& b) {
	double n = 0, d = 0;
	for (size_t i = 0; i < a.size(); ++i) { n += ((double)a[i] - b[i]) * ((double)a[i] - b[i]); d += (double)b[i] * b[i]; }
	return d > 0 ? sqrt(n / d) : sqrt(n);
}

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK failed line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main() {
	const int F = 256, NB = 60;
	AudioEngine engine;
	engine.options.frames_per_buffer = F;
	engine.options.binaural = false;
	engine.options.start_mix_thread = false;
	engine.options.start_streamer_thread = false;
	engine.options.initial_mute_frames = 0;
	engine.options.max_sources = 64;
	std::vector<Reference<SynthStreamer>> made;
	engine.make_streamer = [&](const AudioEngine::AddSourceFromStreamingSoundFileParams& p) {
		Reference<SynthStreamer> s = new SynthStreamer(p.sound_file_path == "a.mp3" ? 32000 : 48000, 40, p.sound_file_path == "a.mp3" ? 330.0f : 520.0f);
		made.push_back(s);
		return Reference<MP3AudioStreamer>(s.ptr());
	};
	try { engine.init(); } catch (glare::Exception& e) { fprintf(stderr, "init failed: %s\n", e.what()); return 3; }
	StreamerThread streamer_thread(&engine);     // pumped by hand: streamOnce() = one pass of StreamerThread::doRun's loop

	orc_engine* orc = orc_create(F, 48000, nullptr, 0);
	struct Mirror { AudioSourceRef src; int oid; size_t fed; float x; int orate; };     // fed: samples of the source's backlog already handed to the oracle
	std::vector<Mirror> ms;
	// The oracle sees what the reference's feeder would pop: everything StreamerThread queued.  Track it per source by watching
	// AudioSource::buffer grow (pushes) and shrink (the mirror's feeder moving samples to the device ring).
	std::vector<std::vector<float>> queued;     // per source: every sample ever queued into AudioSource::buffer, in order
	auto snapshot_pushes = [&](size_t k, size_t before) {
		// samples appended since `before`: read them back from the CircularBuffer copy
		CircularBuffer<float> c = ms[k].src->buffer;
		std::vector<float> all(c.size());
		if (!all.empty()) c.popFrontNItems(all.data(), all.size());
		for (size_t i = before; i < all.size(); ++i) queued[k].push_back(all[i]);
	};
	auto add_streaming = [&](const char* path, bool paused, bool looping, float x) {
		AudioEngine::AddSourceFromStreamingSoundFileParams p;
		p.sound_file_path = path; p.paused = paused; p.looping = looping; p.source_volume = 0.8f; p.global_time = 0.0;
		AudioSourceRef s = engine.addSourceFromStreamingSoundFile(p, Vec4f(x, 0.5f, -3.0f, 1.f));
		Mirror m; m.src = s; m.fed = 0; m.oid = -1; m.x = x;     // the oracle's source is created when the source becomes active:
		ms.push_back(m); queued.emplace_back();                  // a paused AudioSource has no Resonance object (AudioEngine.cpp:887-890)
		snapshot_pushes(ms.size() - 1, 0);     // a source that joins a running stream starts with a copy of the first listener's backlog
		return s;
	};

	CHECK(engine.needNewStreamerForPath("a.mp3", false));
	AudioSourceRef a1 = add_streaming("a.mp3", false, true, 2.0f);
	CHECK(!engine.needNewStreamerForPath("a.mp3", false) && engine.needNewStreamerForPath("a.mp3", true));
	CHECK(engine.unpaused_streams.size() == 1 && engine.stream_to_source_map.size() == 1);

	std::vector<float> got, want, buf(2 * F), obuf(2 * F);
	AudioSourceRef a2, b1;
	for (int b = 0; b < NB; ++b) {
		if (b == 9) {      // a second listener of the same file: shares the decoder, starts in sync
			a2 = add_streaming("a.mp3", false, true, -4.0f);
			CHECK(made.size() == 1 && engine.stream_to_source_map.begin()->second.size() == 2);
			CHECK(a2->buffer.size() == a1->buffer.size());
		}
		if (b == 15) {     // a paused source gets its own decoder and stays silent until it is unpaused
			b1 = add_streaming("b.mp3", true, false, 0.5f);
			CHECK(made.size() == 2 && engine.unpaused_streams.size() == 1 && engine.stream_to_source_map.size() == 2);
		}
		if (b == 25) {
			engine.seekToStartAndUnpauseAudio(*b1);
			queued[2].clear(); ms[2].fed = 0;     // buffer.clear() + stream restart (AudioEngine.cpp:779-797)
			CHECK(made[1]->seeks >= 1);
		}
		if (b == 45) { engine.removeSource(a1); orc_source_destroy(orc, ms[0].oid); ms[0].src = AudioSourceRef(); }
		// StreamerThread pass
		std::vector<size_t> before(ms.size());
		for (size_t k = 0; k < ms.size(); ++k) before[k] = ms[k].src.nonNull() ? ms[k].src->buffer.size() : 0;
		streamer_thread.streamOnce();
		for (size_t k = 0; k < ms.size(); ++k) if (ms[k].src.nonNull()) snapshot_pushes(k, before[k]);
		// the oracle's rings receive what was queued; a decoder-reported rate re-inits the resampler (StreamerThread.cpp:98-105)
		for (size_t k = 0; k < ms.size(); ++k) {
			if (ms[k].src.isNull()) continue;
			AudioSource& s = *ms[k].src;
			if (s.paused) continue;
			if (ms[k].oid < 0) {
				ms[k].oid = orc_source_create(orc, 0);
				orc_source_attach_ring(orc, ms[k].oid, s.sampling_rate); ms[k].orate = s.sampling_rate;
				const float pos[3] = {ms[k].x, 0.5f, -3.0f}; const float vol = 0.8f;
				orc_set_source_positions(orc, 1, &ms[k].oid, pos); orc_set_source_volumes(orc, 1, &ms[k].oid, &vol);
			}
			if (ms[k].orate != s.sampling_rate) { orc_source_attach_ring(orc, ms[k].oid, s.sampling_rate); ms[k].orate = s.sampling_rate; }
			if (queued[k].size() > ms[k].fed) { orc_source_ring_push(orc, ms[k].oid, queued[k].data() + ms[k].fed, queued[k].size() - ms[k].fed); ms[k].fed = queued[k].size(); }
		}
		const bool pushed = engine.mixOneBuffer();
		engine.fillOutputBuffer(buf.data(), F);
		const int ovalid = orc_mix(orc, obuf.data());
		for (float& v : obuf) v = v < -1.f ? -1.f : (v > 1.f ? 1.f : v);
		if (!ovalid) for (float& v : obuf) v = 0.f;
		if (getenv("SBM_TEST_VERBOSE")) { double n = 0, d = 0; for (int i = 0; i < 2 * F; ++i) { n += ((double)buf[i] - obuf[i]) * ((double)buf[i] - obuf[i]); d += (double)obuf[i] * obuf[i]; } printf("b%d err %.3g energy %.3g pushed %d ovalid %d\n", b, d > 0 ? sqrt(n / d) : sqrt(n), d, (int)pushed, ovalid); }
		CHECK(pushed == (ovalid != 0));
		got.insert(got.end(), buf.begin(), buf.end()); want.insert(want.end(), obuf.begin(), obuf.end());
	}
	// the non-looping stream b.mp3 (40 frames of 1152 at 48 kHz = 46080 samples = 180 buffers) is still playing; a.mp3 lost one listener
	CHECK(engine.stream_to_source_map.size() == 2 && engine.stream_to_source_map.begin()->second.size() >= 1);
	engine.removeSource(a2);
	CHECK(engine.unpaused_streams.empty() && engine.stream_to_source_map.size() == 1);
	// Device copies of AudioBuffers follow their users: a one-shot's private buffer goes with it (feeder-finished or removeSource), a
	// buffer that something else still references (here `keep`, as `sound_files` would) stays for the next source.
	{
		const size_t base = engine.numDeviceBuffers();
		AudioBufferRef keep = new AudioBuffer();
		keep->buffer.resize(3 * F); for (size_t i = 0; i < keep->buffer.size(); ++i) keep->buffer[i] = 0.25f;
		for (int round = 0; round < 2; ++round) {
			AudioSourceRef one = new AudioSource();
			one->type = AudioSource::SourceType_NonStreaming; one->looping = false; one->remove_on_finish = true;
			one->shared_buffer = new AudioBuffer();
			one->shared_buffer->buffer.resize(2 * F); for (size_t i = 0; i < one->shared_buffer->buffer.size(); ++i) one->shared_buffer->buffer[i] = 0.5f;
			one->pos = Vec4f(1, 0, -2, 1);
			AudioSourceRef cached = new AudioSource();
			cached->type = AudioSource::SourceType_NonStreaming; cached->looping = true;
			cached->shared_buffer = keep;
			cached->pos = Vec4f(-1, 0, -2, 1);
			engine.addSource(one); engine.addSource(cached);
			CHECK(engine.numDeviceBuffers() == base + 2);
			one->shared_buffer = AudioBufferRef();        // the source lets go of its own reference, as a finished one-shot's owner does
			for (int b = 0; b < 4; ++b) engine.mixOneBuffer();      // 2 buffers of audio, then the feeder finishes it
			CHECK(one->resonance_handle == std::numeric_limits<int>::max());      // INVALID_RESONANCE_HANDLE (AE.cpp:62)
			CHECK(engine.numDeviceBuffers() == base + 1);         // the one-shot's buffer is gone, the cached one is not
			engine.removeSource(cached);
			CHECK(engine.numDeviceBuffers() == base + 1);         // still referenced by `keep`: kept for the next source
		}
		keep = AudioBufferRef();
	}
	const double err = rel_rms(got, want);
	double energy = 0; for (float v : want) energy += (double)v * v;
	printf("streaming host: %d buffers, rel_rms %.3g, energy %.3g, decoders %zu\n", NB, err, energy, made.size());
	engine.shutdown();
	orc_destroy(orc);
	return (err <= 1e-5 && energy > 1.0) ? 0 : 1;
}
