/* sbm_audio.h -- C ABI of the B200 spatial-audio mixer ("sbm" = Substrata B200 Mixer).
 *
 * This library replaces, for Substrata's glare::AudioEngine, the per-buffer body of ResonanceThread::doRun
 * (audio/AudioEngine.cpp:367-537 in the reference tree) together with everything the vendored Resonance Audio
 * library does for it (resonance-audio/resonance_audio/api/resonance_audio_api.h: the subset AudioEngine calls).
 * One sbm_mix() call == "for every active source: gather/resample its sample window, SetPlanarBuffer; then
 * FillInterleavedOutputBuffer".  All arithmetic runs in hand-written sm_100a CUDA kernels; there is no CPU
 * fallback: every entry point that needs the device returns SBM_ERR_CUDA if the device or the kernels are
 * unavailable.
 *
 * Conventions: plain C, opaque handle, int status codes (0 = OK, < 0 = error), no exceptions cross the boundary;
 * pointers are HOST pointers unless the name says `device`.
 *
 * Threading (RA/graph/resonance_audio_api_impl.cc:136-316, RA/utils/lockless_task_queue.cc:43-51): the per-frame
 * setters (sbm_set_source_positions / _volumes / _occlusions, sbm_set_head, sbm_set_master_volume, sbm_source_set_mix_params)
 * may be called from any thread, take no lock and never wait for the mixer or the GPU: each call is queued (lock-free MPSC
 * queue) and takes effect at the next sbm_mix / sbm_mix_begin / sbm_mix_submit, in the order queued.  The structural calls
 * (buffers, source create / attach / destroy, ring pushes) take the engine lock, which is never held while waiting for the
 * GPU.  sbm_mix* itself must be called from one thread at a time (the reference holds AudioEngine::mutex around its
 * per-buffer body, AudioEngine.cpp:367).
 *
 * "file:line" notes name the reference interface each entry point stands in for.
 */
#ifndef SBM_AUDIO_H_
#define SBM_AUDIO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SBM_API __declspec(dllexport)
#else
#define SBM_API __attribute__((visibility("default")))
#endif

typedef struct sbm_engine sbm_engine;

enum {
	SBM_OK = 0,
	SBM_ERR_INVALID_ARG = -1,
	SBM_ERR_CUDA = -2,         /* CUDA runtime / kernel failure; sbm_last_error() has the message */
	SBM_ERR_NO_MEMORY = -3,
	SBM_ERR_UNKNOWN_ID = -4,   /* create / attach / destroy / query calls report unknown ids; the queued setters ignore them silently like
	                              Resonance does (resonance_audio_api_impl.cc:551-554) and count them in sbm_stats.ignored_setter_items */
	SBM_ERR_CAPACITY = -5,     /* more sources / buffers / ring samples than configured */
	SBM_ERR_UNSUPPORTED = -6,  /* rows marked "next" in SURVEY.md 8(f) that are not built */
	SBM_ERR_PEER_TIMEOUT = -7  /* multi-GPU: a peer rank's bus did not arrive in time; the buffer holds this rank's sources only */
};

/* Source kinds.  Values match vraudio::RenderingMode (resonance_audio_api.h:35-45) for sound objects. */
enum {
	SBM_SOURCE_STEREO_PANNING = 0,  /* CreateSoundObjectSource(kStereoPanning): distance attenuation + L/R pan law */
	SBM_SOURCE_BINAURAL_HQ = 3,     /* CreateSoundObjectSource(kBinauralHighQuality), AudioEngine.cpp:857: 3rd-order ambisonic encode + SH-HRIR decode */
	SBM_SOURCE_NONSPATIAL = 100     /* CreateStereoSource(2), AudioEngine.cpp:867: mono duplicated to L/R, volume only */
};

typedef struct sbm_config {
	int frames_per_buffer;      /* Resonance accepts 32..16384 (resonance_audio_api_impl.cc:78-94); this build: 32..1024 */
	int sample_rate_hz;         /* engine rate; AudioEngine uses 48000 (AudioEngine.cpp:580) */
	int device;                 /* CUDA device ordinal */
	int max_sources;            /* slot capacity (sources alive at once) */
	int max_mix_sources;        /* total MixSource entries (AudioEngine.h:61-69) across all sources */
	const float* sh_hrir;       /* planar [16][sh_hrir_taps] SH-domain HRIRs (what CreateShHrirsFromAssets returns,
	                               sh_hrir_creator.cc:60-70); NULL => SBM_SOURCE_BINAURAL_HQ sources are rejected */
	int sh_hrir_taps;           /* 1..1024 (stock SADIE assets: 256) */
} sbm_config;

/* vraudio::CreateResonanceAudioApi(2, frames, rate) -- AudioEngine.cpp:676, resonance_audio_api.h:133-134. */
SBM_API int sbm_create(const sbm_config* cfg, sbm_engine** out_engine);
/* `delete resonance` -- AudioEngine.cpp:837. */
SBM_API void sbm_destroy(sbm_engine* e);
SBM_API const char* sbm_last_error(const sbm_engine* e);

/* ---- PCM buffers: glare::AudioBuffer / SoundFile::buf shared between sources (AudioEngine.h:39-57,122,253) ---- */
/* Uploads n mono f32 samples once; many sources may reference the returned buffer id. */
SBM_API int sbm_buffer_create(sbm_engine* e, const float* samples, size_t n, int* out_buffer);
/* Adopts PCM already resident in device memory (not owned, must outlive the engine): `count` buffers of `n`
 * samples each, buffer k starting at device_base + k*stride_samples.  Ids out_first_buffer .. +count-1.
 * For the fast (TMA) fetch path every buffer must start 16-byte aligned and have >= 3 readable floats after its last
 * sample (i.e. stride_samples >= n rounded up to 4, stride_samples % 4 == 0); other layouts are rendered identically
 * through the generic fetch path. */
SBM_API int sbm_buffer_adopt_device(sbm_engine* e, const float* device_base, size_t n, size_t stride_samples, int count, int* out_first_buffer);
/* Drops the registry's reference to a buffer (`AudioBufferRef` going out of scope, AudioEngine.h:39-45,122): its device memory
 * is freed once the last source that was attached to it is gone and no buffer in flight can still read it.  New attachments
 * to a destroyed buffer are rejected. */
SBM_API int sbm_buffer_destroy(sbm_engine* e, int buffer);

/* ---- sources ---- */
/* CreateSoundObjectSource / CreateStereoSource (AudioEngine.cpp:857,867).  Ids increase monotonically like
 * Resonance's source_id_counter_ (resonance_audio_api_impl.cc:209).  `count` sources are created; ids are
 * out_first_id .. out_first_id+count-1.  New sources are active, volume 1, at the origin. */
SBM_API int sbm_source_create(sbm_engine* e, int kind, int count, int* out_first_id);
/* DestroySource (AudioEngine.cpp:511,908). */
SBM_API int sbm_source_destroy(sbm_engine* e, int n, const int* ids);

/* Feeder attachment = the AudioSource fields the feeder reads (AudioEngine.h:114-131).  Arrays of length n. */
/* SourceType_NonStreaming reading shared_buffer from cur_read_i (AudioEngine.cpp:389-426); resampler.init(rate, engine rate) (:875). */
SBM_API int sbm_source_attach_buffer(sbm_engine* e, int n, const int* ids, const int* buffers, const uint64_t* cur_read_i,
                                     const int* sampling_rate, const uint8_t* looping, const uint8_t* remove_on_finish);
/* SourceType_Streaming fed through the circular buffer (AudioEngine.cpp:459-474). */
SBM_API int sbm_source_attach_ring(sbm_engine* e, int id, int sampling_rate, int capacity_samples);
/* CircularBuffer::pushBackNItems by StreamerThread / ClientUDPHandlerThread (StreamerThread.cpp:53-80, ClientUDPHandlerThread.cpp:239). */
SBM_API int sbm_source_ring_push(sbm_engine* e, int id, const float* samples, size_t n);
/* Batched push for many ring-fed sources at once: packet k = samples[offsets[k] .. offsets[k+1]) goes to source ids[k]
 * (the same id may appear more than once; packets are appended in order).  One pinned staging copy + one H2D + one
 * scatter kernel for the whole call -- what glare::AudioEngine's voice / streaming feeders use instead of one
 * CircularBuffer::pushBackNItems per source (ClientUDPHandlerThread.cpp:239-242, StreamerThread.cpp:33-60).
 * All-or-nothing: SBM_ERR_CAPACITY (some ring full) or SBM_ERR_UNKNOWN_ID leave every ring untouched. */
SBM_API int sbm_source_ring_push_batch(sbm_engine* e, int n, const int* ids, const float* samples, const size_t* offsets /* n + 1 */);
/* Page-locked host memory (cudaHostAlloc) for callers that do not link CUDA: a sbm_source_ring_push_batch payload that
 * lives in it (or in any memory registered with the CUDA driver) goes to the device without the staging copy -- it must
 * then stay unmodified until the buffer SUBMITTED AFTER the push has been collected (synchronous sbm_mix: until the next
 * sbm_mix returns; pipelined: the collect that returns the first buffer submitted after the push -- NOT merely the next
 * collect, which may hand out an older buffer). */
SBM_API void* sbm_host_alloc_pinned(size_t bytes);
SBM_API void sbm_host_free_pinned(void* p);
/* CircularBuffer::size() as of the last collected buffer plus pushes since (StreamerThread.cpp:53 polls it). */
SBM_API int sbm_source_ring_size(sbm_engine* e, int id, uint64_t* out_size);
/* SourceType_Streaming with mix_sources: pitch/doppler mixer (AudioEngine.cpp:435-458). */
SBM_API int sbm_source_attach_mix(sbm_engine* e, int id, int sampling_rate, int nz, const int* buffers, const double* sound_file_i,
                                  const double* source_delta, const float* mix_factor);
/* Batched form for one MixSource per source (doppler_factor pitch shift of a looping buffer). */
SBM_API int sbm_source_attach_mix1(sbm_engine* e, int n, const int* ids, const int* buffers, const double* sound_file_i,
                                   const double* source_delta, const float* mix_factor);
/* mix_sources[z].source_delta / mix_factor updates by the game thread (gui_client/BikePhysics.cpp:712-740). */
SBM_API int sbm_source_set_mix_params(sbm_engine* e, int n, const int* ids, const int* z, const double* source_delta, const float* mix_factor);
/* A streamed source whose decoder reports another sampling rate than the one it was attached with: `source->sampling_rate = hz;
 * source->resampler.init(hz, engine rate)` (StreamerThread.cpp:98-105).  Ring- and buffer-fed sources; the AudioResampler state
 * of the source starts over like AudioResampler::init (AudioResampler.cpp:24-36). */
SBM_API int sbm_source_set_sampling_rate(sbm_engine* e, int id, int sampling_rate);
/* AudioSource::EOF_marker_position (AudioEngine.h:120), AudioSource::looping. */
SBM_API int sbm_source_set_eof_marker(sbm_engine* e, int id, int64_t samples_to_eof);
SBM_API int sbm_source_set_looping(sbm_engine* e, int id, int looping);

/* ---- per-frame setters (AudioEngine.cpp:946-990) ---- */
SBM_API int sbm_set_source_positions(sbm_engine* e, int n, const int* ids, const float* xyz);  /* SetSourcePosition, :955 */
SBM_API int sbm_set_source_volumes(sbm_engine* e, int n, const int* ids, const float* volume); /* SetSourceVolume(volume*mute factor), :965 */
SBM_API int sbm_set_source_occlusions(sbm_engine* e, int n, const int* ids, const float* occ); /* SetSoundObjectOcclusionIntensity, :974 */
SBM_API int sbm_set_head(sbm_engine* e, const float* pos_xyz, const float* quat_xyzw);          /* SetHeadPosition/Rotation, :987-988 */
SBM_API int sbm_set_master_volume(sbm_engine* e, float volume);                                  /* SetMasterVolume, :810 */
/* ---- room effects: early reflections of a shoebox room (AudioEngine.cpp:738-768; RA/graph/graph_manager.cc:106-107,144-151,182-198,
 * RA/graph/reflections_node.cc, RA/dsp/reflections_processor.cc).  Every sound object additionally feeds a mono reflections bus
 * through a ramped gain (room effects gain 1 * direct attenuation * ReflectionProperties::gain); once per buffer the bus goes through six
 * delayed, gain-ramped wall reflections, a first-order ambisonic encode cross-faded whenever the room or the listener moved, the
 * head rotation, and the order-1 SH-HRIR decoder.  The late-reverb send exists too but has gain 0 with Substrata's settings
 * (AudioEngine never calls SetReverbProperties): it contributes zeros and only keeps *out_valid true while sources play. ---- */
/* The order-1 SH-HRIRs the reflections are decoded with: planar [4][taps] (CreateShHrirsFromAssets("...sh_hrir_order_1.wav")).
 * Once, before room effects are first enabled, with no buffer in flight. */
SBM_API int sbm_set_room_sh_hrir(sbm_engine* e, const float* sh_hrir_order1, int taps);
SBM_API int sbm_enable_room_effects(sbm_engine* e, int enable);                                  /* EnableRoomEffects, :742 */
/* SetReflectionProperties, :767 (vraudio::ReflectionProperties, resonance_audio_api.h:73-105): centre, rotation (xyzw) and size of
 * the room, the six wall reflection coefficients (-x +x -y +y -z +z), low-pass cut-off (0 = none) and the reflections gain. */
SBM_API int sbm_set_reflection_properties(sbm_engine* e, const float* room_position, const float* room_rotation_xyzw, const float* room_dimensions,
                                          const float* coefficients6, float cutoff_frequency, float gain);

/* ---- the per-buffer call ---- */
/* SetPlanarBuffer for every active source + FillInterleavedOutputBuffer(2, frames, out) (AudioEngine.cpp:370-526).
 * out_interleaved: 2*frames floats (LRLR...).  *out_valid = what Fill... returns: 0 when no source reached the stereo
 * mixer this buffer (out is zero-filled), else 1. */
SBM_API int sbm_mix(sbm_engine* e, float* out_interleaved, int* out_valid);

/* The same call split for multi-GPU use and for device-resident benchmarking:
 *   sbm_mix_begin   uploads pending setters and launches the per-source kernels on the engine's stream; when it
 *                   returns, the partial bus of this rank is (asynchronously) in sbm_bus_device().
 *   (caller may now sum the bus across ranks, e.g. ncclAllReduce on sbm_stream())
 *   sbm_mix_finish  launches SH decode + stereo mix + interleave, copies 2*frames floats to out_interleaved (host)
 *                   and synchronises.  Pass NULL to leave the result on the device (sbm_out_device()).
 * Bus layout: [19][frames] f32: rows 0-15 ambisonic (ACN), 16-17 the stereo bus (panned + non-spatial sources), 18 the mono
 * reflections send (room effects), followed by 4 floats of flags {ambisonic bus, stereo panner, stereo sources, reflections send
 * fed} that sum across ranks like the audio. */
SBM_API int sbm_mix_begin(sbm_engine* e);
SBM_API int sbm_mix_finish(sbm_engine* e, float* out_interleaved, int* out_valid);
/* Pipelined form of sbm_mix: sbm_mix_submit enqueues one whole buffer (uploads, kernels, the copy of the result into
 * pinned memory) and returns without waiting; sbm_mix_collect waits for the OLDEST submitted buffer and hands it out.
 * At most two buffers may be in flight, so the host's launch latency hides behind the previous buffer's kernels --
 * ResonanceThread::doRun likewise runs ahead of the device callback through its ring (AudioEngine.cpp:540-545).
 * Setters called between two submits apply to the later buffer, like Resonance's task queue.  sbm_poll_finished and
 * the ring capacity seen by sbm_source_ring_push* lag by the buffers not yet collected. */
SBM_API int sbm_mix_submit(sbm_engine* e);
SBM_API int sbm_mix_collect(sbm_engine* e, float* out_interleaved, int* out_valid);

/* Multi-GPU (one engine per GPU, one process per GPU, sources sharded by the caller): the only exchange of
 * the path -- the sum of the ranks' [19][frames] buses -- runs over NVLink peer memory inside the kernels: the CTA of K3
 * that finishes last copies this rank's bus into every rank's gather area and signals, a gather kernel in front of K4 waits
 * and sums in rank order.
 *   sbm_peer_export : allocates this engine's gather area and returns its 64-byte CUDA IPC handle;
 *   sbm_peer_connect: after the ranks exchanged their handles (any host transport; handles + p * handle_stride is rank
 *                     p's), maps the peers.  From then on every sbm_mix / sbm_mix_begin+finish is a collective: all
 *                     ranks must call it for every buffer.  world == 1 is allowed (no peers).
 * Without these calls an engine is single-GPU; a caller may also sum buses itself between sbm_mix_begin and
 * sbm_mix_finish (sbm_bus_device), which is what the CPU-only tests do over gloo. */
SBM_API int sbm_peer_export(sbm_engine* e, void* handle_out, size_t handle_bytes /* >= 64 */);
SBM_API int sbm_peer_connect(sbm_engine* e, int rank, int world, const void* handles, size_t handle_stride);
/* A rank that stalls (debugger, long upload, a failed sbm_mix that skipped a buffer) must not take the others down: the gather
 * kernel gives up after the timeout (default 2000 ms), the buffer is rendered from this rank's sources alone and
 * sbm_mix / sbm_mix_collect return SBM_ERR_PEER_TIMEOUT; the engine stays usable.  After a timeout (or after any rank skipped
 * a buffer) the ranks' exchange counts disagree: quiesce all ranks (nothing in flight), synchronise them on the host, and
 * call sbm_peer_resync on every rank before mixing again. */
SBM_API int sbm_set_peer_timeout_ms(sbm_engine* e, int timeout_ms);
SBM_API int sbm_peer_resync(sbm_engine* e);
SBM_API int sbm_bus_device(sbm_engine* e, float** out_device_ptr, size_t* out_num_floats);
/* The interleaved result (2*frames floats) of the buffer finished or submitted LAST, on the device: the engine keeps one result
 * block per buffer in flight, so the pointer alternates from buffer to buffer -- ask again after every sbm_mix_finish /
 * sbm_mix_submit, and order reads behind that buffer on sbm_stream(). */
SBM_API int sbm_out_device(sbm_engine* e, float** out_device_ptr);
SBM_API void* sbm_stream(sbm_engine* e);   /* cudaStream_t */

/* ---- feeder bookkeeping read-back (AudioEngine.cpp:496-517) ---- */
/* Sources the feeder stopped (non-looping source reached its end / EOF marker): they were destroyed exactly like
 * `resonance->DestroySource` + erase from active_audio_sources.  Returns the number written to ids (<= cap). */
SBM_API int sbm_poll_finished(sbm_engine* e, int* ids, int cap, int* out_n);

typedef struct sbm_source_state {
	uint64_t cur_read_i;      /* AudioSource::cur_read_i */
	int64_t eof_marker;       /* AudioSource::EOF_marker_position */
	uint64_t ring_size;       /* CircularBuffer::size() */
	double mix_pos0;          /* mix_sources[0].sound_file_i */
	float direct_gain;        /* GainNode's current gain */
	float occlusion;          /* OcclusionNode::current_occlusion_ */
	int alive;
} sbm_source_state;
SBM_API int sbm_source_query(sbm_engine* e, int id, sbm_source_state* out);

/* ---- statistics (the reference has none; SURVEY.md section 5 "metrics") ---- */
typedef struct sbm_stats {
	uint64_t buffers_mixed;
	uint64_t kernel_launches;        /* kernels launched by this library so far */
	uint32_t active_sources;
	float last_params_kernel_ms;     /* CUDA-event times of the last buffer on the engine stream (sbm_set_profiling): K1 */
	float last_mix_kernel_ms;        /*   K2 launches only (the HBM-bound gather/mix kernel) */
	float last_finish_kernel_ms;     /*   K3 + whatever the caller put between begin and finish + K4 */
	uint64_t last_algorithmic_bytes; /* PCM bytes the rendered sources had to read: sum of 4 * source samples */
	uint64_t last_h2d_bytes, last_d2h_bytes;
	uint32_t last_rendered_sources;  /* sources that reached a bus in the last buffer */
	uint32_t ignored_setter_items;   /* queued setter items dropped so far because their id was unknown when they were applied */
} sbm_stats;
SBM_API int sbm_get_stats(sbm_engine* e, sbm_stats* out);
/* Enables CUDA-event timing of the kernels inside mix_begin/finish (off by default). */
SBM_API int sbm_set_profiling(sbm_engine* e, int enable);

/* ---- AudioResampler drop-in (audio/AudioResampler.h:23-42) for callers that resample outside the engine.
 * Batched: `count` independent streams, each converting src_rate -> dst_rate with the reference's linear
 * interpolation and carried state.  All pointers are HOST memory; work runs on the GPU. ---- */
typedef struct sbm_resampler sbm_resampler;
SBM_API int sbm_resampler_create(int device, int count, const int* src_rate, int dst_rate, sbm_resampler** out);
SBM_API void sbm_resampler_destroy(sbm_resampler* r);
/* numSrcSamplesNeeded(dest_num_samples) for every stream (AudioResampler.cpp:77-86). */
SBM_API int sbm_resampler_num_src_needed(sbm_resampler* r, size_t n_dst, size_t* out_needed);
/* resample(): stream k reads src + src_offsets[k] .. +needed[k], writes dst + k*n_dst (AudioResampler.cpp:89-113). */
SBM_API int sbm_resampler_resample(sbm_resampler* r, float* dst, size_t n_dst, const float* src, const size_t* src_offsets, size_t src_total);

/* ---- vraudio::Resampler drop-in: Resonance Audio's polyphase sample-rate converter (RA/dsp/resampler.h:27-130), which the AudioEngine
 * path uses to bring the SH-HRIRs to a device rate other than 48 kHz (RA/dsp/sh_hrir_creator.cc:42-55).  Same methods, same
 * argument meaning; audio is planar [channels][frames] in HOST memory; the filtering runs on the GPU and reproduces the reference
 * bit for bit. ---- */
typedef struct sbm_polyphase sbm_polyphase;
SBM_API int sbm_polyphase_create(int device, sbm_polyphase** out);
SBM_API void sbm_polyphase_destroy(sbm_polyphase* p);
SBM_API int sbm_polyphase_rates_supported(int source_hz, int destination_hz);                         /* Resampler::AreSampleRatesSupported: 1 / 0 */
SBM_API int sbm_polyphase_set_rate_and_channels(sbm_polyphase* p, int source_hz, int destination_hz, int num_channels);   /* SetRateAndNumChannels */
SBM_API void sbm_polyphase_reset_state(sbm_polyphase* p);                                              /* ResetState */
SBM_API size_t sbm_polyphase_max_output_length(const sbm_polyphase* p, size_t input_length);           /* GetMaxOutputLength */
SBM_API size_t sbm_polyphase_next_output_length(const sbm_polyphase* p, size_t input_length);          /* GetNextOutputLength */
/* Process: n_out must be sbm_polyphase_next_output_length(n_in). */
SBM_API int sbm_polyphase_process(sbm_polyphase* p, const float* input, size_t n_in, float* output, size_t n_out);
/* CreateShHrirsFromWav's resampling step: SH-HRIRs [channels][taps] at source_hz -> destination_hz.  out == NULL: only *out_taps. */
SBM_API int sbm_resample_sh_hrir(int device, const float* sh_hrir, int channels, int taps, int source_hz, int destination_hz, float* out,
                                 int out_cap_taps, int* out_taps);

/* ---- debugging / test hooks (host-side exact-arithmetic helpers shared with the kernels) ---- */
SBM_API float sbm_debug_ramp_advance(float g, float inc, uint32_t n);   /* value after n float adds g += inc, closed form */
SBM_API float sbm_debug_atan2f(float y, float x);                        /* bit-exact restatement of glibc 2.39 atan2f */
SBM_API const float* sbm_debug_encoder_table(void);                      /* 91*91*15 floats */
SBM_API int sbm_debug_get_ambisonic_bus(sbm_engine* e, float* out16xF);  /* this rank's 16 x frames bus of the last mix */

#ifdef __cplusplus
}
#endif
#endif /* SBM_AUDIO_H_ */
