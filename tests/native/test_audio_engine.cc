// tests/native/test_audio_engine.cc -- drives the glare::AudioEngine mirror (substrata_b200/host/AudioEngine.h) the way
// Substrata's client does: addSource(AudioSourceRef), setHeadTransform every buffer, the mixing loop body, the device
// callback.  Reads a scenario file written by tests/test_gpu_host_engine.py and writes the callback's output, which the
// pytest compares with the oracle.  Usage: test_audio_engine <scenario.bin> <out.bin> [threaded]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>

#include "AudioEngine.h"

using namespace glare;

template <class T> static T rd(FILE* f) { T v; if (fread(&v, sizeof(T), 1, f) != 1) { fprintf(stderr, "short read\n"); exit(2); } return v; }
static void rdv(FILE* f, float* p, size_t n) { if (fread(p, sizeof(float), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

int main(int argc, char** argv) {
	if (argc < 3) return 2;
	const bool threaded = argc > 3 && !strcmp(argv[3], "threaded");
	FILE* f = fopen(argv[1], "rb");
	if (!f) return 2;
	const int n = rd<int>(f), F = rd<int>(f), nb = rd<int>(f), loop_len = rd<int>(f), binaural = rd<int>(f), taps = rd<int>(f);
	AudioEngine engine;
	engine.options.frames_per_buffer = F;
	engine.options.binaural = binaural != 0;
	engine.options.start_mix_thread = threaded;
	engine.options.initial_mute_frames = 0;
	engine.options.max_sources = n + 16;
	if (binaural) { engine.options.sh_hrir.resize((size_t)16 * taps); rdv(f, engine.options.sh_hrir.data(), (size_t)16 * taps); engine.options.sh_hrir_taps = taps; }
	try {
		engine.init();
	} catch (glare::Exception& e) { fprintf(stderr, "init failed: %s\n", e.what()); return 3; }

	std::vector<AudioSourceRef> sources;
	std::vector<std::vector<float>> streams;     // content of the streamed sources
	std::vector<int> stream_src;
	for (int s = 0; s < n; ++s) {
		AudioSourceRef src = new AudioSource();
		const int nonspatial = rd<int>(f);
		const int streaming = rd<int>(f);
		src->sampling_rate = rd<int>(f);
		src->looping = rd<int>(f) != 0;
		src->cur_read_i = (size_t)rd<uint64_t>(f);
		src->volume = rd<float>(f);
		const float x = rd<float>(f), y = rd<float>(f), z = rd<float>(f);
		src->pos = Vec4f(x, y, z, 1.f);
		src->spatial_type = nonspatial ? AudioSource::SourceSpatialType_NonSpatial : AudioSource::SourceSpatialType_Spatial;
		std::vector<float> pcm((size_t)loop_len);
		rdv(f, pcm.data(), (size_t)loop_len);
		if (streaming) {
			// a voice / streamed source: other threads push packets into AudioSource::buffer (ClientUDPHandlerThread.cpp:239-242)
			src->type = AudioSource::SourceType_Streaming;
			streams.push_back(pcm); stream_src.push_back(s);
		} else {
			src->type = AudioSource::SourceType_NonStreaming;
			src->shared_buffer = new AudioBuffer();
			src->shared_buffer->buffer.resize((size_t)loop_len);
			memcpy(src->shared_buffer->buffer.data(), pcm.data(), sizeof(float) * (size_t)loop_len);
		}
		engine.addSource(src);
		sources.push_back(src);
	}
	// error behaviour of addSource (AudioEngine.cpp:884-885)
	bool threw = false;
	try { AudioSourceRef bad = new AudioSource(); bad->sampling_rate = 96000; engine.addSource(bad); } catch (glare::Exception&) { threw = true; }
	if (!threw) { fprintf(stderr, "addSource accepted a 96 kHz source\n"); return 4; }

	std::vector<float> out((size_t)nb * 2 * F);
	for (int b = 0; b < nb; ++b) {
		float hp[3], hq[4];
		rdv(f, hp, 3); rdv(f, hq, 4);
		engine.setHeadTransform(Vec4f(hp[0], hp[1], hp[2], 1.f), Quatf(hq[0], hq[1], hq[2], hq[3]));
		// packets of the streamed sources for this buffer (480 samples per 10 ms at 48 kHz, scaled by the source rate); one
		// buffer in five delivers nothing (underflow -> the mixer zero-pads, AudioEngine.cpp:466-473)
		if (b % 5 != 3) {
			Lock lock(engine.mutex);
			for (size_t k = 0; k < streams.size(); ++k) {
				AudioSource& src = *sources[(size_t)stream_src[k]];
				const size_t m = (size_t)(480 * ((F + 479) / 480) * (double)src.sampling_rate / 48000.0);
				const size_t at = (size_t)b * m;
				if (at + m <= streams[k].size()) src.buffer.pushBackNItems(streams[k].data() + at, m);
			}
		}
		if (!threaded) engine.mixOneBuffer();
		else std::this_thread::sleep_for(std::chrono::milliseconds(20));
		engine.fillOutputBuffer(out.data() + (size_t)b * 2 * F, (size_t)F);
	}
	fclose(f);
	// registry bookkeeping as AudioEngine::test() checks it (AudioEngine.cpp:1162-1466): removeSource empties the sets
	for (AudioSourceRef& s : sources) engine.removeSource(s);
	{
		Lock lock(engine.mutex);
		if (!engine.audio_sources.empty() || !engine.active_audio_sources.empty()) { fprintf(stderr, "registry not empty after removeSource\n"); return 5; }
	}
	engine.shutdown();
	FILE* o = fopen(argv[2], "wb");
	fwrite(out.data(), sizeof(float), out.size(), o);
	fclose(o);
	return 0;
}
