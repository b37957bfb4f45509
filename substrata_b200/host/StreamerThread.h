// StreamerThread.h -- mirror of glare::StreamerThread (audio/StreamerThread.h:28-45): streams decoded audio from the
// MP3AudioStreamers into the AudioSource::buffer of every source that plays them.  Host-side feed logic only; the samples reach
// the device rings in AudioEngine::mixOneBuffer (one batched sbm_source_ring_push_batch per buffer).
#pragma once
#include "AudioEngine.h"

namespace glare {

class StreamerThread : public MessageableThread {
public:
	StreamerThread(AudioEngine* audio_engine);
	virtual void doRun() override;
	virtual void kill() override { die = 1; }
	// One pass of doRun's loop body over all streams (StreamerThread.cpp:31-116); doRun = { streamOnce(); sleep 2 ms } until killed.
	void streamOnce();

	AudioEngine* audio_engine;
	glare::AtomicInt die;

	// Temp buffers
	js::Vector<float, 16> samples;
	js::Vector<float, 16> mono_samples;
};

}  // namespace glare
