// AudioEngine.cpp -- host side of the glare::AudioEngine mirror (see AudioEngine.h).  Follows audio/AudioEngine.cpp of
// the reference tree function by function; "AE.cpp:n" cites its lines.  No sample arithmetic happens here: the only
// per-sample host work is the device-callback clamp (AE.cpp:215-216), which the reference also does on the host.
#include "AudioEngine.h"
#include "StreamerThread.h"

#include <chrono>
#include <limits>

static const int INVALID_RESONANCE_HANDLE = std::numeric_limits<int>::max();   // AE.cpp:62

namespace glare {

// ------------------------------------------------------------------------------------------------------------------
// AudioSource (AE.cpp:68-163)
// ------------------------------------------------------------------------------------------------------------------
AudioSource::AudioSource()
    : resonance_handle(INVALID_RESONANCE_HANDLE), sampling_rate(44100), EOF_marker_position(std::numeric_limits<int64>::max()), cur_read_i(0),
      type(SourceType_NonStreaming), spatial_type(SourceSpatialType_Spatial), paused(false), looping(true), remove_on_finish(false), volume(1.f),
      mute_volume_factor(1.f), mute_change_start_time(-2), mute_change_end_time(-1), mute_vol_fac_start(1.f), mute_vol_fac_end(1.f),
      doppler_factor(1), pos(0, 0, 0, 1), num_occlusions(0), userdata_1(0), smoothed_cur_level(0), feed_prev_dest_dst_coords(-1),
      feed_prev_samples_0_src_coords(-2), device_sampling_rate(0), device_eof_marker(std::numeric_limits<int64>::max()) {}

AudioSource::~AudioSource() {}

void AudioSource::startMuting(double cur_time, double transition_period) {
	if (mute_vol_fac_end != 0) {   // If we have not already started muting, and are not already muted
		mute_change_start_time = cur_time;
		mute_change_end_time = cur_time + transition_period * mute_volume_factor;
		mute_vol_fac_start = mute_volume_factor;
		mute_vol_fac_end = 0;
	}
}

void AudioSource::startUnmuting(double cur_time, double transition_period) {
	if (mute_vol_fac_end != 1) {
		mute_change_start_time = cur_time;
		mute_change_end_time = cur_time + transition_period * (1.f - mute_volume_factor);
		mute_vol_fac_start = mute_volume_factor;
		mute_vol_fac_end = 1;
	}
}

void AudioSource::updateCurrentMuteVolumeFactor(double cur_time) {
	if ((cur_time < mute_change_end_time) && ((mute_change_end_time - mute_change_start_time) > 1.0e-4)) {
		const double frac = (cur_time - mute_change_start_time) / (mute_change_end_time - mute_change_start_time);
		mute_volume_factor = mute_vol_fac_start + (float)frac * (mute_vol_fac_end - mute_vol_fac_start);
	} else {
		mute_volume_factor = mute_vol_fac_end;
	}
}

void AudioSource::setMuteVolumeFactorImmediately(float factor) {
	mute_change_end_time = -1;
	mute_volume_factor = factor;
}

void AudioSource::updateDopplerEffectFactor(const Vec4f& source_linear_vel, const Vec4f& listener_linear_vel, const Vec4f& listener_pos) {
	const Vec4f source_to_listener = listener_pos - pos;
	const Vec4f projected_source_vel = projectOntoDir(source_linear_vel, source_to_listener);
	const Vec4f projected_listener_vel = projectOntoDir(listener_linear_vel, source_to_listener);
	const float v_s = projected_source_vel.length() * Maths::sign(dot(projected_source_vel, source_to_listener));
	const float v_l = projected_listener_vel.length() * Maths::sign(dot(projected_listener_vel, source_to_listener));
	const float c = 343.f;   // Speed of sound in air.
	this->doppler_factor = (c - v_l) / (c - v_s);
}

bool AudioSource::isPlaying() { return !paused && (looping || (EOF_marker_position > 0)); }

float SoundFile::maxVal() const {
	float m = -std::numeric_limits<float>::infinity();
	for (float x : buf->buffer) m = myMax(m, x);
	return m;
}
float SoundFile::minVal() const {
	float m = std::numeric_limits<float>::infinity();
	for (float x : buf->buffer) m = myMin(m, x);
	return m;
}

// ------------------------------------------------------------------------------------------------------------------
// AudioEngine
// ------------------------------------------------------------------------------------------------------------------
AudioEngine::AudioEngine() : engine(nullptr), initialised(false), die(false), frames_mixed(0), sample_rate(48000) { callback_data.engine = this; }

AudioEngine::~AudioEngine() { shutdown(); }

// AE.cpp:576-735 without the OS device: create the mixer (CreateResonanceAudioApi, :676), start the mixing thread (:711-718).
void AudioEngine::init() {
	if (initialised) return;
	sample_rate = (uint32)options.sample_rate;
	sbm_config cfg;
	cfg.frames_per_buffer = options.frames_per_buffer;
	cfg.sample_rate_hz = options.sample_rate;
	cfg.device = options.cuda_device;
	cfg.max_sources = options.max_sources;
	cfg.max_mix_sources = options.max_mix_sources;
	cfg.sh_hrir = options.binaural && !options.sh_hrir.empty() ? options.sh_hrir.data() : nullptr;
	cfg.sh_hrir_taps = options.sh_hrir_taps;
	if (options.binaural && !cfg.sh_hrir) throw glare::Exception("AudioEngine::init: binaural rendering needs the SH-HRIR taps (Options::sh_hrir)");
	const int rc = sbm_create(&cfg, &engine);
	if (rc != SBM_OK) {
		engine = nullptr;
		throw glare::Exception("Failed to initialise the B200 audio mixer (sbm_create returned " + std::to_string(rc) + "); there is no CPU fallback");
	}
	if (!options.sh_hrir_order1.empty() &&
	    sbm_set_room_sh_hrir(engine, options.sh_hrir_order1.data(), options.sh_hrir_order1_taps) != SBM_OK)
		throw glare::Exception(std::string("sbm_set_room_sh_hrir failed: ") + sbm_last_error(engine));
	sbm_enable_room_effects(engine, 0);   // AE.cpp:742 / GUIClient.cpp:6896: room effects off outside parcels
	frames_mixed = 0;
	die = false;
	initialised = true;
	// AE.cpp:711-726: the mixing thread (ResonanceThread in the reference) and the StreamerThread live under thread_manager
	if (options.start_mix_thread) {
		class MixThread : public MessageableThread {
		public:
			explicit MixThread(AudioEngine* e_) : e(e_) {}
			void doRun() override { e->mixThreadMain(); }
			void kill() override { e->die = true; }
			AudioEngine* e;
		};
		thread_manager.addThread(Reference<MessageableThread>(new MixThread(this)));
	}
	if (options.start_streamer_thread) thread_manager.addThread(Reference<MessageableThread>(new StreamerThread(this)));
}

void AudioEngine::shutdown() {
	die = true;
	thread_manager.killThreadsBlocking();   // AE.cpp:816
	{
		Lock lock(mutex);
		active_audio_sources.clear();
		audio_sources.clear();
		source_of_handle.clear();
		unpaused_streams.clear();
		stream_to_source_map.clear();
	}
	if (engine) { sbm_destroy(engine); engine = nullptr; }          // delete resonance, AE.cpp:837
	device_buffers.clear();
	buffers_of_handle.clear();
	initialised = false;
}

// Uploads an AudioBuffer once; sources sharing it (sound_files cache, AE.h:122,253) share the device copy.
int AudioEngine::deviceBufferFor(const AudioBufferRef& buf, int for_handle) {
	DeviceBuffer& d = device_buffers[buf.ptr()];
	if (d.id < 0) {
		int id = -1;
		const int rc = sbm_buffer_create(engine, buf->buffer.data(), buf->buffer.size(), &id);
		if (rc != SBM_OK) { device_buffers.erase(buf.ptr()); throw glare::Exception(std::string("sbm_buffer_create failed: ") + sbm_last_error(engine)); }
		d.id = id; d.keep = buf;
	}
	d.users++;
	buffers_of_handle[for_handle].push_back(buf.ptr());
	return d.id;
}

// A device source has gone (removeSource, or the feeder finished it): its buffers lose a user.
void AudioEngine::releaseDeviceBuffersOf(int handle) {
	auto f = buffers_of_handle.find(handle);
	if (f == buffers_of_handle.end()) return;
	for (const AudioBuffer* b : f->second) {
		auto it = device_buffers.find(b);
		if (it == device_buffers.end()) continue;
		if (--it->second.users <= 0 && it->second.keep->getRefCount() == 1) {
			sbm_buffer_destroy(engine, it->second.id);      // freed on the device once no submitted buffer can still read it
			device_buffers.erase(it);
		}
	}
	buffers_of_handle.erase(f);
}

// AE.cpp:848-876
void AudioEngine::createResonanceObAndResamplerForSource(AudioSource& source_) {
	if (!initialised) return;
	AudioSource* source = &source_;
	int id = -1;
	if (source->spatial_type == AudioSource::SourceSpatialType_Spatial) {
		if (sbm_source_create(engine, options.binaural ? SBM_SOURCE_BINAURAL_HQ : SBM_SOURCE_STEREO_PANNING, 1, &id) != SBM_OK)
			throw glare::Exception(std::string("sbm_source_create failed: ") + sbm_last_error(engine));
		source->resonance_handle = id;
		if (source->pos.isFinite()) {   // Avoid NaN or Inf position coords.
			const float xyz[3] = {source->pos[0], source->pos[1], source->pos[2]};
			sbm_set_source_positions(engine, 1, &id, xyz);
			const float v = source->volume * source->getMuteVolumeFactor();
			sbm_set_source_volumes(engine, 1, &id, &v);
		}
	} else {
		if (sbm_source_create(engine, SBM_SOURCE_NONSPATIAL, 1, &id) != SBM_OK)
			throw glare::Exception(std::string("sbm_source_create failed: ") + sbm_last_error(engine));
		source->resonance_handle = id;
		const float v = source->volume * source->getMuteVolumeFactor();
		sbm_set_source_volumes(engine, 1, &id, &v);
	}
	// Feeder attachment: what ResonanceThread reads from the AudioSource every buffer (AE.cpp:389-474); resampler.init (:875).
	int rc = SBM_OK;
	if (source->type == AudioSource::SourceType_NonStreaming) {
		if (source->shared_buffer.isNull()) throw glare::Exception("AudioSource of type NonStreaming has no shared_buffer");
		const int buf = deviceBufferFor(source->shared_buffer, id);
		const uint64_t cur = source->cur_read_i;
		const uint8_t looping = source->looping, rem = source->remove_on_finish;
		rc = sbm_source_attach_buffer(engine, 1, &id, &buf, &cur, &source->sampling_rate, &looping, &rem);
	} else if (!source->mix_sources.empty()) {
		std::vector<int> bufs; std::vector<double> pos, delta; std::vector<float> mixf;
		for (MixSource& m : source->mix_sources) {
			bufs.push_back(deviceBufferFor(m.soundfile->buf, id));
			pos.push_back(m.sound_file_i); delta.push_back(m.source_delta); mixf.push_back(m.mix_factor);
		}
		rc = sbm_source_attach_mix(engine, id, source->sampling_rate, (int)bufs.size(), bufs.data(), pos.data(), delta.data(), mixf.data());
	} else {
		rc = sbm_source_attach_ring(engine, id, source->sampling_rate, 4 * 48000);
		source->feed_prev_dest_dst_coords = -1; source->feed_prev_samples_0_src_coords = -2;
		source->device_sampling_rate = source->sampling_rate;
	}
	if (rc != SBM_OK) {
		const std::string msg = sbm_last_error(engine);
		sbm_source_destroy(engine, 1, &id);
		releaseDeviceBuffersOf(id);
		source->resonance_handle = INVALID_RESONANCE_HANDLE;
		throw glare::Exception("attaching audio data to the source failed: " + msg);
	}
	if (source->EOF_marker_position != std::numeric_limits<int64>::max()) sbm_source_set_eof_marker(engine, id, source->EOF_marker_position);
	source->device_eof_marker = source->EOF_marker_position;
	source->resampler.init(/*src rate=*/source->sampling_rate, (int)this->sample_rate, options.cuda_device);   // AE.cpp:875
	if (source->num_occlusions != 0) sbm_set_source_occlusions(engine, 1, &id, &source->num_occlusions);
}

// AE.cpp:879-897
void AudioEngine::addSource(AudioSourceRef source) {
	if (!initialised) return;
	if (source->sampling_rate < 8000 || source->sampling_rate > 48000)
		throw glare::Exception("Unsupported sampling rate for audio source: " + std::to_string(source->sampling_rate));
	Lock lock(mutex);   // the mixing thread must not run between creating the source and registering it
	if (!source->paused) createResonanceObAndResamplerForSource(*source);
	audio_sources.insert(source);
	if (!source->paused) {
		active_audio_sources.insert(source);
		source_of_handle[source->resonance_handle] = source;
	}
}

// AE.cpp:900-943 (the MP3 streamer maps are out of scope)
void AudioEngine::removeSource(AudioSourceRef source) {
	if (!initialised) return;
	Lock lock(mutex);
	if (source->resonance_handle != INVALID_RESONANCE_HANDLE) {
		sbm_source_destroy(engine, 1, &source->resonance_handle);
		releaseDeviceBuffersOf(source->resonance_handle);
		source_of_handle.erase(source->resonance_handle);
		source->resonance_handle = INVALID_RESONANCE_HANDLE;
	}
	audio_sources.erase(source);
	active_audio_sources.erase(source);
	// Remove the source from the stream -> sources map; the last listener of a stream frees its decoder (AE.cpp:915-940).
	Reference<MP3AudioStreamer> streamer_to_remove;
	for (auto it = stream_to_source_map.begin(); it != stream_to_source_map.end(); ++it) {
		it->second.erase(source);
		if (it->second.empty()) streamer_to_remove = it->first;
	}
	if (streamer_to_remove) {
		for (auto it = unpaused_streams.begin(); it != unpaused_streams.end(); ++it)
			if (it->second == streamer_to_remove) { unpaused_streams.erase(it); break; }
		stream_to_source_map.erase(streamer_to_remove);
	}
}

// AE.cpp:1081-1140
AudioSourceRef AudioEngine::addSourceFromStreamingSoundFile(AddSourceFromStreamingSoundFileParams& params, const Vec4f& pos) {
	Lock lock(mutex);
	AudioSourceRef source = new AudioSource();
	source->type = AudioSource::SourceType_Streaming;
	source->remove_on_finish = false;
	source->pos = pos;
	source->volume = params.source_volume;
	source->looping = params.looping;
	source->paused = params.paused;
	// A new decoder if the source is paused (a script may unpause it at any time, independently of the other listeners of the
	// file) or if nobody streams this file yet; otherwise the running one is shared.
	auto res = unpaused_streams.find(params.sound_file_path);
	if ((res == unpaused_streams.end()) || params.paused) {
		if (!make_streamer) throw glare::Exception("AudioEngine::make_streamer is not set: the mp3 decoder is supplied by the owner of the engine");
		Reference<MP3AudioStreamer> streamer = make_streamer(params);
		if (streamer.isNull()) throw glare::Exception("make_streamer returned no decoder for " + params.sound_file_path);
		if (!params.paused) {   // if not paused, the stream can be shared with other sources
			streamer->seekToApproxTimeWrapped(params.global_time);
			unpaused_streams.insert(std::make_pair(params.sound_file_path, streamer));
		}
		stream_to_source_map[streamer].insert(source);
	} else {
		Reference<MP3AudioStreamer> streamer = res->second;
		// another source already plays this stream: start from a copy of its backlog, so that both consume the stream in step
		auto first_source_it = stream_to_source_map[streamer].begin();
		if (first_source_it != stream_to_source_map[streamer].end()) {
			AudioSourceRef first_source = *first_source_it;
			source->buffer = first_source->buffer;
		}
		stream_to_source_map[streamer].insert(source);
	}
	addSource(source);
	return source;
}

bool AudioEngine::needNewStreamerForPath(const std::string& sound_file_path, bool new_source_paused) {   // AE.cpp:1142-1146
	auto res = unpaused_streams.find(sound_file_path);
	return (res == unpaused_streams.end()) || new_source_paused;
}

// AE.cpp:946-956.  The notification only queues the new value; mixOneBuffer sends all queued ones as ONE sbm_set_source_positions
// call (they take effect at the next buffer either way: Resonance queues its setters too, resonance_audio_api_impl.cc:305-317).
void AudioEngine::sourcePositionUpdated(AudioSource& source) {
	if (!initialised) return;
	if (!source.pos.isFinite()) return;
	if (source.resonance_handle != INVALID_RESONANCE_HANDLE) {
		std::lock_guard<std::mutex> lk(pending_mutex);
		pend_pos_ids.push_back(source.resonance_handle);
		pend_pos.push_back(source.pos[0]); pend_pos.push_back(source.pos[1]); pend_pos.push_back(source.pos[2]);
	}
}

// AE.cpp:959-966
void AudioEngine::sourceVolumeUpdated(AudioSource& source) {
	if (!initialised) return;
	if (source.resonance_handle != INVALID_RESONANCE_HANDLE) {
		std::lock_guard<std::mutex> lk(pending_mutex);
		pend_vol_ids.push_back(source.resonance_handle);
		pend_vol.push_back(source.volume * source.getMuteVolumeFactor());
	}
}

// AE.cpp:969-975
void AudioEngine::sourceNumOcclusionsUpdated(AudioSource& source) {
	if (!initialised) return;
	if (source.resonance_handle != INVALID_RESONANCE_HANDLE) {
		std::lock_guard<std::mutex> lk(pending_mutex);
		pend_occ_ids.push_back(source.resonance_handle);
		pend_occ.push_back(source.num_occlusions);
	}
}

// The client's per-frame loop over all sources (GUIClient.cpp:6979-7034) in batched form.
void AudioEngine::sourcesPositionUpdated(AudioSource* const* sources, size_t n) {
	if (!initialised) return;
	std::lock_guard<std::mutex> lk(pending_mutex);
	for (size_t i = 0; i < n; ++i) {
		const AudioSource& s = *sources[i];
		if (s.resonance_handle == INVALID_RESONANCE_HANDLE || !s.pos.isFinite()) continue;
		pend_pos_ids.push_back(s.resonance_handle);
		pend_pos.push_back(s.pos[0]); pend_pos.push_back(s.pos[1]); pend_pos.push_back(s.pos[2]);
	}
}
void AudioEngine::sourcesVolumeUpdated(AudioSource* const* sources, size_t n) {
	if (!initialised) return;
	std::lock_guard<std::mutex> lk(pending_mutex);
	for (size_t i = 0; i < n; ++i) {
		const AudioSource& s = *sources[i];
		if (s.resonance_handle == INVALID_RESONANCE_HANDLE) continue;
		pend_vol_ids.push_back(s.resonance_handle);
		pend_vol.push_back(s.volume * s.getMuteVolumeFactor());
	}
}
void AudioEngine::sourcesNumOcclusionsUpdated(AudioSource* const* sources, size_t n) {
	if (!initialised) return;
	std::lock_guard<std::mutex> lk(pending_mutex);
	for (size_t i = 0; i < n; ++i) {
		const AudioSource& s = *sources[i];
		if (s.resonance_handle == INVALID_RESONANCE_HANDLE) continue;
		pend_occ_ids.push_back(s.resonance_handle);
		pend_occ.push_back(s.num_occlusions);
	}
}

// Everything the notifications queued since the last buffer: one C-ABI call per kind.
void AudioEngine::flushPendingUpdates() {
	std::lock_guard<std::mutex> lk(pending_mutex);
	if (!pend_pos_ids.empty()) sbm_set_source_positions(engine, (int)pend_pos_ids.size(), pend_pos_ids.data(), pend_pos.data());
	if (!pend_vol_ids.empty()) sbm_set_source_volumes(engine, (int)pend_vol_ids.size(), pend_vol_ids.data(), pend_vol.data());
	if (!pend_occ_ids.empty()) sbm_set_source_occlusions(engine, (int)pend_occ_ids.size(), pend_occ_ids.data(), pend_occ.data());
	pend_pos_ids.clear(); pend_pos.clear(); pend_vol_ids.clear(); pend_vol.clear(); pend_occ_ids.clear(); pend_occ.clear();
}

// AE.cpp:978-990
void AudioEngine::setHeadTransform(const Vec4f& head_pos, const Quatf& head_rot) {
	if (!initialised) return;
	if (head_pos.isFinite() && head_rot.v.isFinite()) {
		const float p[3] = {head_pos[0], head_pos[1], head_pos[2]};
		const float q[4] = {head_rot.v[0], head_rot.v[1], head_rot.v[2], head_rot.v[3]};
		sbm_set_head(engine, p, q);
	}
}

void AudioEngine::setMasterVolume(float volume) {   // AE.cpp:806-811
	if (!initialised) return;
	sbm_set_master_volume(engine, volume);
}

void AudioEngine::setRoomEffectsEnabled(bool enabled) {   // AE.cpp:738-743
	if (!initialised) return;
	if (enabled && options.sh_hrir_order1.empty()) throw glare::Exception("room effects need the order-1 SH-HRIR taps (Options::sh_hrir_order1)");
	if (sbm_enable_room_effects(engine, enabled ? 1 : 0) != SBM_OK) throw glare::Exception(std::string("sbm_enable_room_effects failed: ") + sbm_last_error(engine));
}

// AE.cpp:746-768: the parcel's bounding box as Resonance's shoebox room, with Substrata's fixed wall coefficients and gain.
void AudioEngine::setCurentRoomDimensions(const js::AABBox& room_aabb) {
	if (!initialised) return;
	const float dims[3] = {room_aabb.axisLength(0), room_aabb.axisLength(1), room_aabb.axisLength(2)};
	const Vec4f c = room_aabb.centroid();
	const float pos[3] = {c[0], c[1], c[2]};
	const float rot[4] = {0.f, 0.f, 0.f, 1.f};      // vraudio::ReflectionProperties(): room_rotation stays the identity
	float coeffs[6];
	for (int i = 0; i < 6; ++i) coeffs[i] = 0.8f;
	coeffs[5] = 0.2f;
	coeffs[2] = 0.3f;
	sbm_set_reflection_properties(engine, pos, rot, dims, coeffs, /*cutoff_frequency=*/0.f, /*gain=*/0.7f);
}

void AudioEngine::insertSoundFile(const std::string& sound_file_path, SoundFileRef sound) { sound_files.insert(std::make_pair(sound_file_path, sound)); }   // AE.cpp:999-1002

SoundFileRef AudioEngine::getOrLoadSoundFile(const std::string& sound_file_path) {   // AE.cpp:1005-1019, cache part
	auto res = sound_files.find(sound_file_path);
	if (res == sound_files.end()) throw glare::Exception("sound file not in the cache and file decoding is out of scope here: " + sound_file_path);
	return res->second;
}

// AE.cpp:1022-1052
void AudioEngine::playOneShotSound(const std::string& sound_file_path, const Vec4f& pos) {
	if (!initialised) return;
	if (!pos.isFinite()) return;
	try {
		SoundFileRef sound = getOrLoadSoundFile(sound_file_path);
		AudioSourceRef source = new AudioSource();
		source->type = AudioSource::SourceType_NonStreaming;
		source->sampling_rate = (int)sound->sample_rate;
		source->shared_buffer = sound->buf;
		source->paused = false;
		source->looping = false;
		source->remove_on_finish = true;
		source->pos = pos;
		addSource(source);
	} catch (glare::Exception&) {
		// the reference prints a warning and carries on
	}
}

// AE.cpp:770-803: (re)start a paused or finished source from the beginning.
void AudioEngine::seekToStartAndUnpauseAudio(AudioSource& source) {
	if (!initialised) return;
	Lock lock(mutex);
	AudioSourceRef ref(&source);
	source.cur_read_i = 0;
	source.paused = false;
	source.EOF_marker_position = std::numeric_limits<int64>::max();
	source.buffer.clear();
	for (MixSource& m : source.mix_sources) m.sound_file_i = 0;
	// The device-side source holds the read index / ring / EOF state: recreate it (a fresh Resonance object in the reference only
	// when there was none; re-creating is equivalent, the old one's filters start from silence either way after the seek).
	if (source.resonance_handle != INVALID_RESONANCE_HANDLE) {
		sbm_source_destroy(engine, 1, &source.resonance_handle);
		releaseDeviceBuffersOf(source.resonance_handle);
		source_of_handle.erase(source.resonance_handle);
		source.resonance_handle = INVALID_RESONANCE_HANDLE;
	}
	createResonanceObAndResamplerForSource(source);
	// If there is a streamer for this source, then restart the stream.
	for (auto it = stream_to_source_map.begin(); it != stream_to_source_map.end(); ++it)
		if (it->second.count(ref) != 0) it->first->seekToBeginningOfFile();
	active_audio_sources.insert(ref);
	source_of_handle[source.resonance_handle] = ref;
}

// Moves exactly the samples the next buffer will consume from the source's host CircularBuffer to its device ring
// (what `source->buffer.popFrontNItems(...)` does in AE.cpp:459-474), so that buffer.size() keeps the meaning the
// StreamerThread relies on (StreamerThread.cpp:53).  numSrcSamplesNeeded is a pure function of two integer counters.
void AudioEngine::feedStreamingSource(AudioSource& source) {
	const int64 F = options.frames_per_buffer;
	if (source.sampling_rate != source.device_sampling_rate) {
		// StreamerThread found the stream's real rate (StreamerThread.cpp:98-105: source->resampler.init): same on the device
		sbm_source_set_sampling_rate(engine, source.resonance_handle, source.sampling_rate);
		source.device_sampling_rate = source.sampling_rate;
		source.feed_prev_dest_dst_coords = -1; source.feed_prev_samples_0_src_coords = -2;
	}
	if (source.EOF_marker_position != source.device_eof_marker) {
		// StreamerThread placed the EOF marker behind the samples it queued (StreamerThread.cpp:108-109)
		sbm_source_set_eof_marker(engine, source.resonance_handle, source.EOF_marker_position);
		source.device_eof_marker = source.EOF_marker_position;
	}
	size_t n_src = (size_t)F;
	if (source.sampling_rate != (int)sample_rate) {
		const int64 max_dest = source.feed_prev_dest_dst_coords + F;
		const double max_dest_src = (double)max_dest * (double)source.sampling_rate / (double)sample_rate;
		n_src = (size_t)ceil(max_dest_src) - (size_t)(source.feed_prev_samples_0_src_coords + 1);   // AudioResampler.cpp:77-86
		source.feed_prev_dest_dst_coords += F;
		source.feed_prev_samples_0_src_coords += (int64)n_src;
	}
	const size_t have = myMin(n_src, source.buffer.size());
	if (have) {
		// appended to this buffer's batch: all streaming sources cross PCIe in ONE sbm_source_ring_push_batch call
		const size_t at = feed_payload.size();
		feed_payload.resize(at + have);
		source.buffer.popFrontNItems(feed_payload.data() + at, have);
		feed_ids.push_back(source.resonance_handle);
		feed_offsets.push_back(feed_payload.size());
	}
	// AE.cpp:496: the marker counts down by the samples each buffer asks for -- on the device (K1) and here alike
	if (source.EOF_marker_position != std::numeric_limits<int64>::max()) {
		source.EOF_marker_position -= (int64)n_src;
		source.device_eof_marker = source.EOF_marker_position;
	}
}

// The body of ResonanceThread::doRun's inner loop (AE.cpp:367-545).
bool AudioEngine::mixOneBuffer() {
	if (!initialised) return false;
	const size_t F = (size_t)options.frames_per_buffer;
	std::vector<float> out(2 * F);
	int valid = 0;
	{
		Lock lock(mutex);
		feed_ids.clear(); feed_payload.clear(); feed_offsets.assign(1, 0);
		for (auto it = active_audio_sources.begin(); it != active_audio_sources.end(); ++it) {
			AudioSource& source = **it;
			if (source.resonance_handle == INVALID_RESONANCE_HANDLE) continue;
			if (source.type == AudioSource::SourceType_Streaming) {
				if (source.mix_sources.empty()) feedStreamingSource(source);
				else {
					// the game thread edits source_delta / mix_factor in place under `mutex` (BikePhysics.cpp:712-740): collected for one call
					for (size_t z = 0; z < source.mix_sources.size(); ++z) {
						mixp_ids.push_back(source.resonance_handle); mixp_z.push_back((int)z);
						mixp_delta.push_back(source.mix_sources[z].source_delta); mixp_factor.push_back(source.mix_sources[z].mix_factor);
					}
				}
			}
		}
		if (!mixp_ids.empty()) {
			sbm_source_set_mix_params(engine, (int)mixp_ids.size(), mixp_ids.data(), mixp_z.data(), mixp_delta.data(), mixp_factor.data());
			mixp_ids.clear(); mixp_z.clear(); mixp_delta.clear(); mixp_factor.clear();
		}
		flushPendingUpdates();
		if (!feed_ids.empty() && sbm_source_ring_push_batch(engine, (int)feed_ids.size(), feed_ids.data(), feed_payload.data(), feed_offsets.data()) != SBM_OK)
			throw glare::Exception(std::string("sbm_source_ring_push_batch failed: ") + sbm_last_error(engine));
		const int rc = sbm_mix(engine, out.data(), &valid);
		if (rc != SBM_OK) throw glare::Exception(std::string("sbm_mix failed: ") + sbm_last_error(engine));
		// Sources the feeder stopped (AE.cpp:496-517): already destroyed on the device; mirror the registry bookkeeping.
		int ids[256], n = 0;
		do {
			sbm_poll_finished(engine, ids, 256, &n);
			for (int i = 0; i < n; ++i) {
				auto f = source_of_handle.find(ids[i]);
				if (f == source_of_handle.end()) continue;
				AudioSourceRef s = f->second;
				s->resonance_handle = INVALID_RESONANCE_HANDLE;
				releaseDeviceBuffersOf(ids[i]);
				active_audio_sources.erase(s);
				if (s->remove_on_finish) audio_sources.erase(s);
				source_of_handle.erase(f);
			}
		} while (n == 256);
	}
	frames_mixed += (int64)F;
	if (!valid) return false;                                           // AE.cpp:535-536
	// AE.cpp:527-536: `if(buffers_processed * frames_per_buffer < 48000) filled_valid_buffer = false;` -- a buffer of the first
	// second is treated exactly like an invalid one: nothing is pushed and the fill loop breaks (and sleeps 1 ms), so streamed
	// sources are not drained faster than real time during the mute.
	if (frames_mixed < (int64)options.initial_mute_frames) return false;
	{
		Lock lock(callback_data.buffer_mutex);
		callback_data.buffer.pushBackNItems(out.data(), 2 * F);         // AE.cpp:540-545
	}
	return true;
}

// ResonanceThread::doRun (AE.cpp:345-555): keep 4 buffers queued ahead of the device.
void AudioEngine::mixThreadMain() {
	const size_t F = (size_t)options.frames_per_buffer;
	try {
		while (!die) {
			for (;;) {
				size_t queued;
				{ Lock lock(callback_data.buffer_mutex); queued = callback_data.buffer.size(); }
				if (die || queued >= 4 * 2 * F) break;                  // AE.cpp:363
				if (!mixOneBuffer()) break;
			}
			std::this_thread::sleep_for(std::chrono::milliseconds(1));     // AE.cpp:548
		}
	} catch (glare::Exception&) {
		// AE.cpp:551-554: the thread exits on an exception
	}
}

// miniAudioCallBack (AE.cpp:191-226)
void AudioEngine::fillOutputBuffer(float* output_buffer_f, size_t frame_count) {
	const size_t num_samples_needed = frame_count * 2;   // We always use stereo output so each frame is 2 samples.
	Lock lock(callback_data.buffer_mutex);
	const size_t num_samples_to_dequeue = myMin(num_samples_needed, callback_data.buffer.size());
	callback_data.buffer.popFrontNItems(output_buffer_f, num_samples_to_dequeue);
	for (size_t z = 0; z < num_samples_to_dequeue; ++z) output_buffer_f[z] = myClamp(output_buffer_f[z], -1.f, 1.f);
	for (size_t i = num_samples_to_dequeue; i < num_samples_needed; ++i) output_buffer_f[i] = 0.f;   // Pad with zeroes
}

void AudioEngine::syncSourceState(AudioSource& source) {
	if (!initialised || source.resonance_handle == INVALID_RESONANCE_HANDLE) return;
	sbm_source_state st;
	if (sbm_source_query(engine, source.resonance_handle, &st) != SBM_OK) return;
	source.cur_read_i = (size_t)st.cur_read_i;
	source.EOF_marker_position = st.eof_marker;
	if (!source.mix_sources.empty()) source.mix_sources[0].sound_file_i = st.mix_pos0;
}

}  // namespace glare
