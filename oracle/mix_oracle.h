/* oracle/mix_oracle.h -- TEST INFRASTRUCTURE ONLY (see mix_oracle.c).
 *
 * CPU restatement ("port") of the reference mix path.  The entry points mirror include/sbm_audio.h so
 * that tests drive the product, this oracle and the compiled reference (oracle/ref_driver.cc) with one
 * scenario.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 */
#ifndef MIX_ORACLE_H_
#define MIX_ORACLE_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_engine orc_engine;

/* sh_hrir: planar [16][taps] floats (the order-3 SADIE SH-HRIRs as the reference decoder receives them);
 * may be NULL if no HRTF source is ever created. */
orc_engine* orc_create(int frames_per_buffer, int sample_rate_hz, const float* sh_hrir, int taps);
void orc_destroy(orc_engine* e);
/* 0 = float bus (reference arithmetic, sources summed in id order); 1 = double bus ("fp64-bus oracle"). */
void orc_set_bus_precision(orc_engine* e, int use_double);

int  orc_buffer_create(orc_engine* e, const float* samples, size_t n);
int  orc_source_create(orc_engine* e, int kind);  /* 0 stereo-pan object, 3 HQ binaural object, 100 stereo */
int  orc_sources_create(orc_engine* e, int kind, int count);   /* batched orc_source_create: returns the first id */
void orc_sources_attach_buffer(orc_engine* e, int n, const int* ids, const int* bufs, const uint64_t* cur_read_i, const int* rates,
                               const uint8_t* looping, const uint8_t* remove_on_finish);
void orc_sources_attach_mix1(orc_engine* e, int n, const int* ids, const int* bufs, const double* sound_file_i, const double* source_delta, const float* mix_factor);
void orc_source_attach_buffer(orc_engine* e, int id, int buf, uint64_t cur_read_i, int sampling_rate, int looping, int remove_on_finish);
void orc_source_attach_ring(orc_engine* e, int id, int sampling_rate);
void orc_source_ring_push(orc_engine* e, int id, const float* samples, size_t n);
void orc_source_attach_mix(orc_engine* e, int id, int sampling_rate, int nz, const int* bufs, const double* sound_file_i,
                           const double* source_delta, const float* mix_factor);
void orc_source_set_mix_params(orc_engine* e, int id, int z, double source_delta, float mix_factor);
void orc_source_set_eof_marker(orc_engine* e, int id, int64_t samples_to_eof);
void orc_source_set_looping(orc_engine* e, int id, int looping);
void orc_set_source_positions(orc_engine* e, int n, const int* ids, const float* xyz);
void orc_set_source_volumes(orc_engine* e, int n, const int* ids, const float* vol);
void orc_set_source_occlusions(orc_engine* e, int n, const int* ids, const float* occ);
void orc_set_head(orc_engine* e, const float* pos, const float* quat_xyzw);
void orc_set_master_volume(orc_engine* e, float v);
void orc_source_destroy(orc_engine* e, int id);
/* Room effects as Substrata drives them (AudioEngine.cpp:738-768): early reflections in a shoebox room; off by default. */
void orc_set_room_hrir(orc_engine* e, const float* sh_hrir_order1 /* [4][taps] */, int taps);
void orc_enable_room_effects(orc_engine* e, int enable);
void orc_set_reflection_properties(orc_engine* e, const float* room_position, const float* room_rotation_xyzw, const float* room_dimensions,
                                   const float* coefficients6, float cutoff_frequency, float gain);
int  orc_mix(orc_engine* e, float* out_interleaved);
int  orc_get_ambisonic_bus(orc_engine* e, float* out16xF);
int  orc_poll_finished(orc_engine* e, int* ids, int cap);
int  orc_source_query(orc_engine* e, int id, uint64_t* cur_read_i, int64_t* eof_marker, uint64_t* ring_size, double* mix_pos0);

/* ---- stage functions (exposed for the known-answer tests restated from Resonance's *_test.cc) ---- */
float orc_log_distance_attenuation(const float* listener, const float* source, float min_d, float max_d);
float orc_linear_distance_attenuation(const float* listener, const float* source, float min_d, float max_d);
void  orc_update_attenuations(float master, float refl_gain, float reverb_gain, const float* listener, const float* source,
                              float gain, float room_effects_gain, int rolloff, float min_d, float max_d, float user_att, float* out4);
void  orc_stereo_pan_gains(float azimuth_rad, float elevation_rad, float* out2);
void  orc_relative_direction(const float* from_pos, const float* from_quat_xyzw, const float* to_pos, float* out3);
void  orc_spherical_from_world(const float* world3, float* az_el);
void  orc_encoding_coeffs(float azimuth_rad, float elevation_rad, float* out16);
const float* orc_encoder_table(void);  /* 91*91*15 */
float orc_occlusion_filter_coefficient(float directivity, float occlusion);
float orc_directivity(float alpha, float order, float azimuth_rad, float elevation_rad);
/* GainProcessor::ApplyGain.  state[0] = current gain, state[1] = initialised (0/1). */
void  orc_gain_apply(float* state, float target, const float* in, float* out, size_t n, int accumulate);
/* MonoPoleFilter::Filter.  Returns 0 when bypassed (output untouched). */
int   orc_mono_pole(float coefficient, float* prev_out, const float* in, float* out, size_t n);
/* AudioResampler: state = {src_rate, dst_rate, prev_dest_dst_coords, prev_samples_0_src_coords} + prev[2]. */
typedef struct { int src_rate, dst_rate; int64_t prev_dest_dst_coords, prev_samples_0_src_coords; float prev_samples[2]; } orc_resampler;
void   orc_resampler_init(orc_resampler* r, int src_rate, int dst_rate);
size_t orc_resampler_num_src_needed(const orc_resampler* r, size_t n_dst);
void   orc_resampler_resample(orc_resampler* r, float* dst, size_t n_dst, const float* src, size_t n_src);
/* SH-domain binaural decode of one 16-channel buffer (direct-form FIR + tail carry). sh_hrir: [channels][taps]; bus: [channels][frames]; history: [channels][taps-1]. */
void  orc_sh_decode(const float* sh_hrir, int channels, int taps, float* history, const float* bus, int frames, float* out_l, float* out_r);
/* Resonance's polyphase resampler (RA/dsp/resampler.cc), SURVEY.md 8f row 4: planar [channels][frames] in and out. */
typedef struct orc_polyphase orc_polyphase;
orc_polyphase* orc_polyphase_create(void);
void   orc_polyphase_destroy(orc_polyphase* p);
int    orc_polyphase_rates_supported(int source_hz, int destination_hz);
void   orc_polyphase_set_rate_and_channels(orc_polyphase* p, int source_hz, int destination_hz, int num_channels);
void   orc_polyphase_reset_state(orc_polyphase* p);
size_t orc_polyphase_max_output_length(const orc_polyphase* p, size_t input_length);
size_t orc_polyphase_next_output_length(const orc_polyphase* p, size_t input_length);
const float* orc_polyphase_coeffs(const orc_polyphase* p, size_t* coeffs_per_phase, size_t* up_rate);
void   orc_polyphase_process(orc_polyphase* p, const float* input, size_t n_in, float* output, size_t n_out);
#ifdef __cplusplus
}
#endif
#endif
