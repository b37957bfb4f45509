// MP3AudioStreamer.h -- the codec boundary of the streaming feed (audio/MP3AudioStreamer.h in the reference tree): what
// StreamerThread and AudioEngine call on a stream decoder (StreamerThread.cpp:60,80; AudioEngine.cpp:797,1106-1112).  The mp3
// decoder itself (minimp3) is outside the accelerated path, so this is an abstract interface: inside
// Substrata the real class implements it; the tests plug in a synthetic stream.
#pragma once
#ifndef SBM_USE_GLARE_CORE
#include "glare_core_min.h"
#endif

namespace glare {

class MP3AudioStreamer : public ThreadSafeRefCounted {
public:
	virtual ~MP3AudioStreamer() {}
	// Decodes the next frame (an mp3 frame is 1152 sample-frames): interleaved samples, channel count (1 or 2) and sampling
	// rate (0 = unknown).  Returns true when the end of the file was reached.
	virtual bool decodeFrame(js::Vector<float, 16>& samples_out, int& num_channels_out, int& sample_freq_hz_out) = 0;
	virtual void seekToBeginningOfFile() = 0;
	// Seeks to global_time modulo the stream's length, so that everyone who opens the stream at the same global time hears
	// the same position.
	virtual void seekToApproxTimeWrapped(double global_time) = 0;
};

}  // namespace glare
