// AudioEngine.h -- host-side mirror of glare::AudioEngine / glare::AudioSource (audio/AudioEngine.h:39-261 in the
// reference tree): same class and member names, same argument meaning, same error behaviour (init failure and bad
// sampling rates throw glare::Exception; everything else is a silent no-op until initialised).  What changed is
// behind it: where the reference's ResonanceThread gathers every source's window on the CPU and hands it to Resonance
// Audio (AudioEngine.cpp:367-537), this class keeps the sources' PCM resident on the GPU and calls sbm_mix() once per
// buffer (include/sbm_audio.h); `resonance_handle` holds the sbm source id.
//
// Not carried over (out of the hot path): the OS audio device (miniaudio / RtAudio) -- the device
// callback body is exposed as fillOutputBuffer() for whoever owns the device --, the mp3 / audio file DECODERS (streams come in
// through the MP3AudioStreamer interface, getOrLoadSoundFile only consults the cache).  The streaming feed itself is here:
// addSourceFromStreamingSoundFile, needNewStreamerForPath, the stream maps, sync-start of sources that share a stream, and
// StreamerThread (StreamerThread.h) -- all under thread_manager like the reference's.
#pragma once
#ifndef SBM_USE_GLARE_CORE
#include "glare_core_min.h"
#endif
#include <functional>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sbm_audio.h"
#include "AudioResampler.h"
#include "MP3AudioStreamer.h"

namespace glare {

class AudioEngine;

struct AudioBuffer : public ThreadSafeRefCounted {
	js::Vector<float, 16> buffer;
};
typedef Reference<AudioBuffer> AudioBufferRef;
typedef VRef<AudioBuffer> AudioBufferVRef;

struct SoundFile : public ThreadSafeRefCounted {
	SoundFile() : buf(new AudioBuffer()), num_channels(1), sample_rate(48000) {}
	float maxVal() const;
	float minVal() const;
	AudioBufferVRef buf;
	uint32 num_channels;
	uint32 sample_rate;  // in hz
};
typedef Reference<SoundFile> SoundFileRef;

struct MixSource {
	MixSource() : sound_file_i(0), source_delta(1), mix_factor(1) {}
	Reference<SoundFile> soundfile;
	double sound_file_i;  // Current (floating-point) index into soundfile buffer.  Advanced on the GPU; refreshed by AudioEngine::syncSourceState().
	double source_delta;  // Every sample, we advance this much in the soundfile buffer.
	float mix_factor;     // Volume mixing factor.
};

class AudioSource : public ThreadSafeRefCounted {
public:
	GLARE_ALIGNED_16_NEW_DELETE

	enum SourceType {
		SourceType_NonStreaming,  // Audio is played from an audio buffer fully present in RAM (shared_buffer).
		SourceType_Streaming      // Audio data is streamed into the circular buffer, or mixed from mix_sources.
	};
	enum SourceSpatialType { SourceSpatialType_Spatial, SourceSpatialType_NonSpatial };

	AudioSource();
	~AudioSource();

	void startMuting(double cur_time, double transition_period);
	void startUnmuting(double cur_time, double transition_period);
	void updateCurrentMuteVolumeFactor(double cur_time);
	void setMuteVolumeFactorImmediately(float factor);
	inline float getMuteVolumeFactor() const { return mute_volume_factor; }
	void updateDopplerEffectFactor(const Vec4f& source_linear_vel, const Vec4f& listener_linear_vel, const Vec4f& listener_pos);
	bool isPlaying();

	int resonance_handle;  // Set in AudioEngine::addSource(): the sbm source id (INT_MAX = none).
	int sampling_rate;
	CircularBuffer<float> buffer;  // Read from front, enqueue to back.  Used for type SourceType_Streaming only.
	int64 EOF_marker_position;     // Number of samples still to play in buffer before EOF is reached.
	AudioBufferRef shared_buffer;  // Used for SourceType_NonStreaming type only.
	size_t cur_read_i;             // Read index in shared_buffer at addSource(); refreshed by AudioEngine::syncSourceState().
	std::vector<MixSource> mix_sources;
	SourceType type;
	SourceSpatialType spatial_type;
	bool paused;
	bool looping;
	bool remove_on_finish;
	float volume;

private:
	float mute_volume_factor;
	double mute_change_start_time;
	double mute_change_end_time;
	float mute_vol_fac_start;
	float mute_vol_fac_end;

public:
	float doppler_factor;
	Vec4f pos;
	float num_occlusions;
	uint32 userdata_1;
	std::string debugname;
	float smoothed_cur_level;

	AudioResampler resampler;   // AudioEngine.h:153.  (Sources are resampled inside the mix kernels; kept for callers such as StreamerThread that re-init it.)

private:
	friend class AudioEngine;
	// Host mirror of the AudioResampler counters of a ring-fed source (AudioResampler.h:37-39): how many source samples
	// the next buffer consumes is a pure function of these, so the feeder can move exactly that many to the device ring.
	int64 feed_prev_dest_dst_coords, feed_prev_samples_0_src_coords;
	int device_sampling_rate;      // the rate the device side was last told (StreamerThread may change sampling_rate mid-stream)
	int64 device_eof_marker;       // the EOF marker the device side holds, as far as the host knows
};
typedef Reference<AudioSource> AudioSourceRef;

struct AudioCallbackData {
	Mutex buffer_mutex;             // protects buffer
	CircularBuffer<float> buffer GUARDED_BY(buffer_mutex);
	AudioEngine* engine;
};

class AudioEngine {
public:
	AudioEngine();
	~AudioEngine();

	// Extra knobs the reference hard-codes (AudioEngine.cpp:580-586,857) or gets from the device.  Set before init().
	struct Options {
		Options() : frames_per_buffer(256), sample_rate(48000), cuda_device(0), max_sources(65536), max_mix_sources(4096),
		            binaural(true), start_mix_thread(true), start_streamer_thread(true), initial_mute_frames(48000) {}
		int frames_per_buffer;   // device period (256; 1024 on Emscripten)
		int sample_rate;
		int cuda_device;
		int max_sources, max_mix_sources;
		bool binaural;           // true: kBinauralHighQuality for spatial sources as the reference; false: kStereoPanning
		bool start_mix_thread;   // false: the owner calls mixOneBuffer() itself (tests, offline rendering)
		bool start_streamer_thread;   // false: the owner pumps the streams itself (StreamerThread::streamOnce)
		int initial_mute_frames; // the reference discards the first second of output (AudioEngine.cpp:532-533)
		std::vector<float> sh_hrir;  // planar [16][taps] SH-HRIRs (CreateShHrirsFromAssets); required when binaural
		int sh_hrir_taps = 0;
		std::vector<float> sh_hrir_order1;  // planar [4][taps] order-1 SH-HRIRs: the reflections' decoder; required for room effects
		int sh_hrir_order1_taps = 0;
	};
	Options options;

	void init();       // throws glare::Exception when no B200 / the library cannot be initialised (no CPU fallback)
	void shutdown();
	bool isInitialised() const { return initialised; }

	void addSource(AudioSourceRef source);
	void removeSource(AudioSourceRef source);

	// sound_data_source may be null.  If non-null, read from it, otherwise read from disk from sound_file_path  (AudioEngine.h:193-206)
	struct AddSourceFromStreamingSoundFileParams {
		AddSourceFromStreamingSoundFileParams() : source_volume(1.f), global_time(0.0), looping(true), paused(false) {}
		std::string sound_file_path;
		Reference<SharedMemMappedFile> mem_mapped_sound_file;
		Reference<MP3AudioStreamerDataSource> sound_data_source;
		float source_volume;
		double global_time;
		bool looping;
		bool paused;
	};
	// AudioEngine.cpp:1081-1140: a new streaming source; sources that play the same (unpaused) file share one decoder, and a
	// source that joins a running stream starts from a copy of the first listener's backlog so that all of them stay in sync.
	AudioSourceRef addSourceFromStreamingSoundFile(AddSourceFromStreamingSoundFileParams& params, const Vec4f& pos);
	bool needNewStreamerForPath(const std::string& sound_file_path, bool new_source_paused);   // AudioEngine.cpp:1142-1146
	// Where the reference constructs `new MP3AudioStreamer(data source | mem-mapped file)` (AudioEngine.cpp:1106-1110) this mirror
	// asks the owner for the decoder: the codec is outside the accelerated path.  Must be set before the first streaming source.
	std::function<Reference<MP3AudioStreamer>(const AddSourceFromStreamingSoundFileParams&)> make_streamer;

	// Batched forms of the three per-frame notifications below for the O(N) per-frame loop of the client
	// (gui_client/GUIClient.cpp:6979-7034): one C-ABI call for all of them instead of one per source.
	void sourcesPositionUpdated(AudioSource* const* sources, size_t n);
	void sourcesVolumeUpdated(AudioSource* const* sources, size_t n);
	void sourcesNumOcclusionsUpdated(AudioSource* const* sources, size_t n);

	void sourcePositionUpdated(AudioSource& source);
	void sourceVolumeUpdated(AudioSource& source);
	void sourceNumOcclusionsUpdated(AudioSource& source);
	void setHeadTransform(const Vec4f& head_pos, const Quatf& head_rot);
	void playOneShotSound(const std::string& sound_file_path, const Vec4f& pos);
	void setRoomEffectsEnabled(bool enabled);                      // early reflections on the GPU (needs Options::sh_hrir_order1)
	void setCurentRoomDimensions(const js::AABBox& room_aabb);
	uint32 getSampleRate() const { return sample_rate; }
	void seekToStartAndUnpauseAudio(AudioSource& source);
	void setMasterVolume(float volume);
	void insertSoundFile(const std::string& sound_file_path, SoundFileRef sound);
	SoundFileRef getOrLoadSoundFile(const std::string& sound_file_path);   // cache only: file decoding is out of scope

	// ---- what the reference does in ResonanceThread::doRun / miniAudioCallBack, exposed for the owner of the device ----
	// One pass of the mixing loop body (AudioEngine.cpp:367-545): feed streamed sources, mix one buffer on the GPU,
	// apply feeder bookkeeping (finished sources), push 2 * frames samples to the callback ring.  Returns false when
	// nothing reached the mixer (nothing pushed), like FillInterleavedOutputBuffer.
	bool mixOneBuffer();
	size_t numDeviceBuffers() const { return device_buffers.size(); }   // AudioBuffers that currently have a device copy
	// miniAudioCallBack (AudioEngine.cpp:191-226): pop, clamp to [-1, 1], zero-pad.
	void fillOutputBuffer(float* output_buffer, size_t frame_count);
	// Refreshes cur_read_i / EOF_marker_position / mix_sources[].sound_file_i of `source` from the device.
	void syncSourceState(AudioSource& source);
	sbm_engine* nativeEngine() { return engine; }

private:
	void createResonanceObAndResamplerForSource(AudioSource& source);
	int deviceBufferFor(const AudioBufferRef& buf, int for_handle);
	void feedStreamingSource(AudioSource& source);
	void mixThreadMain();

	void flushPendingUpdates();

	sbm_engine* engine;
	AudioCallbackData callback_data;
	bool initialised;
	std::atomic<bool> die;
	// The single-source notifications (sourcePositionUpdated ...) only append here; mixOneBuffer sends each kind as ONE batched
	// sbm_set_* call per buffer.  They take effect at the same point as before: the next buffer.
	std::mutex pending_mutex;
	std::vector<int> pend_pos_ids, pend_vol_ids, pend_occ_ids;
	std::vector<float> pend_pos, pend_vol, pend_occ;
	std::vector<int> mixp_ids, mixp_z;
	std::vector<double> mixp_delta;
	std::vector<float> mixp_factor;
	int64 frames_mixed;
	// AudioBuffer -> its device copy.  `keep` holds the host buffer alive (so the key cannot be recycled for another buffer), `users` counts
	// the live device sources reading it.  The device copy is dropped (sbm_buffer_destroy) when the last such source has gone AND nobody
	// but this table still references the host buffer -- a SoundFile that stays in `sound_files` keeps its device copy for the next source.
	struct DeviceBuffer { int id = -1; AudioBufferRef keep; int users = 0; };
	std::map<const AudioBuffer*, DeviceBuffer> device_buffers;
	std::map<int, std::vector<const AudioBuffer*> > buffers_of_handle;   // source handle -> the AudioBuffers it reads
	void releaseDeviceBuffersOf(int handle);
	std::map<int, AudioSourceRef> source_of_handle;
	std::vector<float> temp_buf;
	// this buffer's streaming feed, pushed to the device rings in one batched call (sbm_source_ring_push_batch)
	std::vector<int> feed_ids;
	std::vector<float> feed_payload;
	std::vector<size_t> feed_offsets;

public:
	Mutex mutex;  // Guards access to audio_sources
	std::set<AudioSourceRef> active_audio_sources GUARDED_BY(mutex);
	std::set<AudioSourceRef> audio_sources GUARDED_BY(mutex);

	ThreadManager thread_manager;  // Manages: the mixing thread (ResonanceThread in the reference), StreamerThread

	std::map<std::string, SoundFileRef> sound_files;

	std::map<std::string, Reference<MP3AudioStreamer>> unpaused_streams;  // Map from mp3 file path to MP3AudioStreamer for that mp3.

	std::map<Reference<MP3AudioStreamer>, std::set<AudioSourceRef>> stream_to_source_map;  // Map from MP3AudioStreamer to audio sources playing that stream.

private:
	uint32 sample_rate;
};

}  // namespace glare
