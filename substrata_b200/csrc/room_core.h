// room_core.h -- room effects behind the per-source reflections send: Resonance's ReflectionsNode / ReflectionsProcessor /
// FoaRotator (RA/graph/reflections_node.cc, RA/dsp/reflections_processor.cc, RA/dsp/delay_filter.cc,
// RA/utils/buffer_crossfader.cc, RA/ambisonics/foa_rotator.cc) as `__host__ __device__` functions, so that the room kernel
// (device_cuda.cu: one CTA, one thread per frame) and the host emulation of the no-GPU tests evaluate the same statements.
// O(1) in the number of sources: input = the mono reflections bus (bus row kRowRefl, summed over CTAs and ranks), output =
// a first-order ambisonic buffer [4][frames] (ACN: W Y Z X) that the order-1 SH-HRIR decode turns into stereo.
//
// Split: room_plan() -- ONE thread: what the node does this buffer (process or not, crossfade or not, the six gain
// processors' ramps, the rotation mode); room_frame() -- every frame independently.  The host decides what
// ReflectionsNode::Update decides (engine.cc: room_update_host) and ships it as RoomParams.
#pragma once
#include "k2_core.h"

namespace sbm {

constexpr int kRoomSurfaces = 6;

struct RoomParams {             // host-written, one per buffer (rides in the upload block)
	uint32_t enabled;           // room effects on: the send mixers are unmuted
	uint32_t updated;           // ReflectionsNode::Update ran ReflectionsProcessor::Update this buffer: new targets below, crossfade
	uint32_t tgt_delay[kRoomSurfaces];   // target_reflections_[i].delay_time_seconds * sample_rate, truncated, min'ed with max_delay
	float tgt_gain[kRoomSurfaces];       // target_reflections_[i].magnitude
	float lp_coef;              // low-pass coefficient (0 = bypass; Substrata never sets a cut-off)
	uint32_t frames_to_process_on_empty;
	float head_q_inv[4];        // conjugate of the head rotation, xyzw
	uint32_t delay_size;        // delay line length = frames + 2 s
	uint32_t pad[3];
};

struct RoomState {              // device-resident, written by the room kernel only
	float gp_gain[kRoomSurfaces]; uint32_t gp_init[kRoomSurfaces];   // gain_processors_
	uint32_t delays[kRoomSurfaces]; float gains[kRoomSurfaces];      // delays_, gains_
	uint32_t tgt_delay[kRoomSurfaces]; float tgt_gain[kRoomSurfaces];
	float lp_coef, lp_prev;
	uint32_t crossfade;
	uint32_t frames_to_process_on_empty, frames_processed_on_empty;
	uint32_t write_cursor;
	float rot_cur[4];           // FoaRotator::current_rotation_
};

struct RoomPlan {
	uint32_t process;           // 0: the node returns no buffer this cycle
	uint32_t passes;            // 2: crossfade from pass 0 (current gains / delays) to pass 1 (targets)
	uint32_t silent_input;      // process on silence (tail flush)
	RampDesc g[2][kRoomSurfaces];
	uint32_t on[2][kRoomSurfaces];      // surface contributes in that pass
	uint32_t delay[2][kRoomSurfaces];
	uint32_t read_base;         // write cursor after this buffer's insert
	uint32_t rot_mode;          // 0 none, 1 constant (qa), 2 slerp qa -> qb
	float qa[4], qb[4];
	float lp_coef, lp_prev;
	uint32_t delay_size;
};

SBM_HD void rc_cross(const float* a, const float* b, float* o) {
	o[0] = fsub(fmul(a[1], b[2]), fmul(a[2], b[1]));
	o[1] = fsub(fmul(a[2], b[0]), fmul(a[0], b[2]));
	o[2] = fsub(fmul(a[0], b[1]), fmul(a[1], b[0]));
}
SBM_HD void rc_quat_mul(const float* a, const float* b, float* o) {
	o[3] = fsub(fsub(fsub(fmul(a[3], b[3]), fmul(a[0], b[0])), fmul(a[1], b[1])), fmul(a[2], b[2]));
	o[0] = fsub(fadd(fadd(fmul(a[3], b[0]), fmul(a[0], b[3])), fmul(a[1], b[2])), fmul(a[2], b[1]));
	o[1] = fsub(fadd(fadd(fmul(a[3], b[1]), fmul(a[1], b[3])), fmul(a[2], b[0])), fmul(a[0], b[2]));
	o[2] = fsub(fadd(fadd(fmul(a[3], b[2]), fmul(a[2], b[3])), fmul(a[0], b[1])), fmul(a[1], b[0]));
}
// WorldRotation::AngularDifferenceRad (RA/base/misc_math.h:69-72; Eigen AngleAxis.h:170-191).  Compared against the 1-degree
// rotation quantum: exact float operations and the bit-exact atan2f, so that the branch falls like the reference's.
SBM_HD float rc_angular_difference(const float* a, const float* b) {
	const float n2 = fadd(fadd(fadd(fmul(a[0], a[0]), fmul(a[1], a[1])), fmul(a[2], a[2])), fmul(a[3], a[3]));
	float inv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
	if (n2 > 0.0f) { inv[0] = fdiv(-a[0], n2); inv[1] = fdiv(-a[1], n2); inv[2] = fdiv(-a[2], n2); inv[3] = fdiv(a[3], n2); }
	float d[4];
	rc_quat_mul(inv, b, d);
	const float n = fsqrt(fadd(fadd(fmul(d[0], d[0]), fmul(d[1], d[1])), fmul(d[2], d[2])));
	return n != 0.0f ? fmul(2.0f, exact_atan2f(n, fabsf(d[3]))) : 0.0f;
}
// QuaternionBase::slerp (Eigen Quaternion.h:782-810).  Continuous in its inputs: the device's acosf / sinf serve.
SBM_HD void rc_slerp(const float* a, float t, const float* b, float* o) {
	const float one = 1.0f - 1.1920929e-07f;
	const float d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
	const float absd = fabsf(d);
	float s0, s1;
	if (absd >= one) { s0 = 1.0f - t; s1 = t; }
	else {
		const float theta = acosf(absd), st = sinf(theta);
		s0 = sinf((1.0f - t) * theta) / st;
		s1 = sinf(t * theta) / st;
	}
	if (d < 0.0f) s1 = -s1;
	for (int i = 0; i < 4; ++i) o[i] = s0 * a[i] + s1 * b[i];
}
SBM_HD void rc_rotate(const float* q, const float* v, float* o) {   // q * v (Eigen Quaternion.h:531-541)
	float uv[3], c[3];
	rc_cross(q, v, uv);
	uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
	rc_cross(q, uv, c);
	for (int i = 0; i < 3; ++i) o[i] = (v[i] + q[3] * uv[i]) + c[i];
}

// One pass of ReflectionsProcessor::ApplyReflections (reflections_processor.cc:134-174), planning part: the gain processors.
SBM_HD void rc_plan_pass(RoomState* st, uint32_t F, RampDesc* g, uint32_t* on, uint32_t* delay) {
	for (int s = 0; s < kRoomSurfaces; ++s) {
		delay[s] = st->delays[s];
		const bool zero_gain = near_zero(st->gains[s]) && near_zero(st->gp_gain[s]);
		if (!zero_gain) {
			g[s] = plan_gain(&st->gp_gain[s], st->gp_init[s] != 0, st->gains[s], F, /*accumulate=*/false);
			st->gp_init[s] = 1;
			on[s] = 1;
		} else {
			st->gp_gain[s] = 0.0f; st->gp_init[s] = 1;      // gain_processors_[i].Reset(0.0f)
			g[s].konst = 0.0f; g[s].start = 0.0f; g[s].inc = 0.0f; g[s].n = 0;
			on[s] = 0;
		}
	}
}

// ReflectionsNode::AudioProcess + ReflectionsProcessor::Process, everything that is decided once per buffer.
// have_input: the reflections mixer produced a buffer (room effects enabled and some sound object was fed, on any rank).
SBM_HD void room_plan(const RoomParams& p, RoomState* st, bool have_input, uint32_t F, RoomPlan* pl) {
	if (p.updated) {
		// ReflectionsProcessor::Update (reflections_processor.cc:74-97)
		for (int s = 0; s < kRoomSurfaces; ++s) { st->tgt_delay[s] = p.tgt_delay[s]; st->tgt_gain[s] = p.tgt_gain[s]; }
		st->lp_coef = p.lp_coef;
		st->frames_to_process_on_empty = p.frames_to_process_on_empty;
		st->crossfade = 1;
	}
	pl->process = 1; pl->silent_input = 0;
	if (!have_input) {
		// reflections_node.cc:88-98: silence until the delay line has played out
		if (st->frames_processed_on_empty < st->frames_to_process_on_empty) { st->frames_processed_on_empty += F; pl->silent_input = 1; }
		else { pl->process = 0; return; }
	} else st->frames_processed_on_empty = 0;
	pl->delay_size = p.delay_size;
	pl->lp_coef = st->lp_coef; pl->lp_prev = st->lp_prev;
	st->write_cursor = (st->write_cursor + F) % p.delay_size;      // DelayFilter::InsertData (the frames themselves: room_insert)
	pl->read_base = st->write_cursor;
	if (st->crossfade) {
		pl->passes = 2;
		rc_plan_pass(st, F, pl->g[0], pl->on[0], pl->delay[0]);
		for (int s = 0; s < kRoomSurfaces; ++s) { st->delays[s] = st->tgt_delay[s]; st->gains[s] = st->tgt_gain[s]; }   // UpdateGainsAndDelays (:123-132)
		rc_plan_pass(st, F, pl->g[1], pl->on[1], pl->delay[1]);
		st->crossfade = 0;
	} else {
		pl->passes = 1;
		rc_plan_pass(st, F, pl->g[0], pl->on[0], pl->delay[0]);
	}
	// FoaRotator::Process with the inverse head rotation (foa_rotator.cc:30-78)
	const float ident[4] = {0.0f, 0.0f, 0.0f, 1.0f};
	const float quant = (float)(3.14159265358979323846 / 180.0);
	for (int i = 0; i < 4; ++i) { pl->qa[i] = st->rot_cur[i]; pl->qb[i] = p.head_q_inv[i]; }
	if (rc_angular_difference(st->rot_cur, ident) < quant && rc_angular_difference(p.head_q_inv, ident) < quant) pl->rot_mode = 0;
	else if (rc_angular_difference(st->rot_cur, p.head_q_inv) < quant) pl->rot_mode = 1;
	else { pl->rot_mode = 2; for (int i = 0; i < 4; ++i) st->rot_cur[i] = p.head_q_inv[i]; }
}

// The low-passed input sample i is written to the delay line at (old write cursor + i); with lp_coef >= 1e-6 the caller runs
// the recurrence serially first (MonoPoleFilter::Filter, mono_pole_filter.cc:34-57) -- room_lowpass -- else the input passes.
SBM_HD uint32_t room_insert_index(const RoomPlan& pl, uint32_t F, uint32_t i) {
	return (pl.read_base + pl.delay_size - F + i) % pl.delay_size;
}
SBM_HD float room_lowpass(float coef, float prev, const float* in, float* out, uint32_t F) {
	for (uint32_t n = 0; n < F; ++n) { prev = lpf_step(coef, prev, in[n]); out[n] = prev; }
	return prev;
}

// Frame i of the node's output: six delayed copies through their gain ramps, encoded to first order (W += all; Y +- the x-walls;
// Z -+ floor / ceiling; X +- the z-walls), cross-faded between the two passes, rotated into the head frame.
SBM_HD void room_frame(const RoomPlan& pl, const float* line, uint32_t F, uint32_t i, float* w, float* y, float* z, float* x) {
	float acc[2][4];
	for (uint32_t ps = 0; ps < pl.passes; ++ps) {
		float W = 0.0f, Y = 0.0f, Z = 0.0f, X = 0.0f;
		for (int s = 0; s < kRoomSurfaces; ++s) {
			if (!pl.on[ps][s]) continue;
			// DelayFilter::GetDelayedData (delay_filter.cc:140-176): frames_per_buffer + delay samples behind the write cursor
			const uint32_t at = (pl.read_base + 2u * pl.delay_size - pl.delay[ps][s] - F + i) % pl.delay_size;
			const float v = fmul(line[at], gain_at(pl.g[ps][s], i));
			W = fadd(W, v);
			switch (s) {
				case 0: Y = fadd(Y, v); break;      // left wall
				case 1: Y = fsub(Y, v); break;      // right wall
				case 2: Z = fsub(Z, v); break;      // floor
				case 3: Z = fadd(Z, v); break;      // ceiling
				case 4: X = fadd(X, v); break;      // front wall
				default: X = fsub(X, v); break;     // back wall
			}
		}
		acc[ps][0] = W; acc[ps][1] = Y; acc[ps][2] = Z; acc[ps][3] = X;
	}
	float W, Y, Z, X;
	if (pl.passes == 2) {
		// BufferCrossfader::ApplyLinearCrossfade (buffer_crossfader.cc:40-65): target fades in, current fades out
		const float fin = fdiv((float)i, (float)F), fout = fsub(1.0f, fin);
		W = fadd(fmul(fin, acc[1][0]), fmul(fout, acc[0][0])); Y = fadd(fmul(fin, acc[1][1]), fmul(fout, acc[0][1]));
		Z = fadd(fmul(fin, acc[1][2]), fmul(fout, acc[0][2])); X = fadd(fmul(fin, acc[1][3]), fmul(fout, acc[0][3]));
	} else { W = acc[0][0]; Y = acc[0][1]; Z = acc[0][2]; X = acc[0][3]; }
	if (pl.rot_mode) {
		float q[4];
		if (pl.rot_mode == 1) { for (int k = 0; k < 4; ++k) q[k] = pl.qa[k]; }
		else {
			// slerp per 32-frame chunk, factor = end of the chunk / frames
			const uint32_t c0 = i & ~31u;
			const uint32_t dur = F - c0 < 32u ? F - c0 : 32u;
			rc_slerp(pl.qa, (float)(c0 + dur) / (float)F, pl.qb, q);
		}
		// audio (x, y, z) -> world (-y, z, -x); rotate; world -> audio (-z, -x, y)  (RA/base/misc_math.h:91-115)
		const float wv[3] = {-Y, Z, -X};
		float r[3];
		rc_rotate(q, wv, r);
		X = -r[2]; Y = -r[0]; Z = r[1];
	}
	*w = W; *y = Y; *z = Z; *x = X;
}

}  // namespace sbm
