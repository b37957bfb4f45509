// StreamerThread.cpp -- see StreamerThread.h.  Follows audio/StreamerThread.cpp:27-120 of the reference tree.
#include "StreamerThread.h"

#include <chrono>
#include <limits>

namespace glare {

StreamerThread::StreamerThread(AudioEngine* audio_engine_) : audio_engine(audio_engine_), die(0) {}

void StreamerThread::streamOnce() {
	Lock lock(audio_engine->mutex);
	for (auto it = audio_engine->stream_to_source_map.begin(); it != audio_engine->stream_to_source_map.end(); ++it) {
		MP3AudioStreamer* streamer = it->first.ptr();
		const std::set<AudioSourceRef>& sources_playing_stream = it->second;
		if (sources_playing_stream.empty()) continue;
		// the smallest backlog over all sources playing this stream decides how much to decode (StreamerThread.cpp:45-48)
		size_t min_src_buffer_size = std::numeric_limits<size_t>::max();
		for (const AudioSourceRef& src : sources_playing_stream) min_src_buffer_size = myMin(min_src_buffer_size, src->buffer.size());
		// Up to 4 mp3 frames (1152 sample-frames each) per pass refill the 4096-sample backlog (92 ms at 44.1 kHz): :50-55
		const int max_num_iters = 4;
		int iter = 0;
		while (min_src_buffer_size < 4096) {
			int num_channels = 0, sample_freq_hz = 0;
			const bool is_EOF = streamer->decodeFrame(samples, num_channels, sample_freq_hz);
			if (num_channels == 1) mono_samples = samples;
			else if (num_channels == 2) {   // stereo -> mono: the mean of the two channels (:66-75)
				const size_t num_frames = samples.size() / 2;
				if (samples.size() == num_frames * 2) {
					mono_samples.resizeNoCopy(num_frames);
					for (size_t i = 0; i < num_frames; ++i) mono_samples[i] = (samples[i * 2 + 0] + samples[i * 2 + 1]) * 0.5f;
				}
			} else mono_samples.resize(0);
			if (is_EOF) streamer->seekToBeginningOfFile();
			// not a valid stream: no samples ever arrive; do not loop forever (:86-89)
			iter++;
			if (iter > max_num_iters) break;
			for (const AudioSourceRef& src : sources_playing_stream) {
				AudioSource* source = src.ptr();
				if (source->buffer.size() < 48000)   // Avoid filling the buffer too much for paused sources.
					source->buffer.pushBackNItems(mono_samples.data(), mono_samples.size());
				if (sample_freq_hz != 0 && source->sampling_rate != sample_freq_hz) {
					// the stream's real rate differs from the default: re-init the resampler (:98-105).  The device side follows in
					// the next AudioEngine::mixOneBuffer (sbm_source_set_sampling_rate).
					source->sampling_rate = sample_freq_hz;
					source->resampler.init(/*src rate=*/sample_freq_hz, /*dest rate=*/(int)audio_engine->getSampleRate());
				}
				if (is_EOF) source->EOF_marker_position = (int64)source->buffer.size();   // after the samples just enqueued
			}
			min_src_buffer_size += mono_samples.size();
		}
	}
}

void StreamerThread::doRun() {
	while (die == 0) {
		streamOnce();
		std::this_thread::sleep_for(std::chrono::milliseconds(2));   // PlatformUtils::Sleep(2), :118
	}
}

}  // namespace glare
