// k1_core.h -- the per-source, per-buffer scalar work ("K1"): feeder bookkeeping + Resonance's per-source
// parameter logic, condensed into one WorkRec for the mix kernel.  One call per slot per buffer.
//
// Follows, in order: ResonanceThread::doRun's per-source body (AE.cpp:375-517), AudioResampler::numSrcSamplesNeeded /
// resample's coordinate bookkeeping (AR.cpp:77-113), UpdateAttenuationParameters (RA/dsp/distance_attenuation.cc:82-119),
// GainNode::AudioProcess (RA/graph/gain_node.cc:38-80), OcclusionNode::AudioProcess (RA/graph/occlusion_node.cc:53-99),
// StereoMixingPannerNode / AmbisonicMixingEncoderNode (RA/graph/stereo_mixing_panner_node.cc:33-63,
// RA/graph/ambisonic_mixing_encoder_node.cc:36-67) and GainMixer's processor bookkeeping (RA/dsp/gain_mixer.cc:53-113).
// Host + device: see sbm_core.h.
#pragma once
#include "sbm_core.h"

namespace sbm {

SBM_HD int finished_push(DevFlags* f) {
#if defined(__CUDA_ARCH__)
	return atomicAdd(&f->n_finished, 1);
#else
	return f->n_finished++;
#endif
}
// Bits returned by update_source (OR-reduced per block by the caller into DevFlags::*_used[parity]).
constexpr uint32_t kUsedPan = 1, kUsedSh = 2, kUsedStereo = 4, kUsedRefl = 8;

// StereoPanner gains straight from the relative direction r: with az = atan2(-r.x, -r.z) and el = atan2(r.y, |r.xz|),
// cos(pi/2 - az) cos(el) = sin(az) cos(el) = -r.x / |r|, so  gL = 0.5 (1 - r.x/|r|),  gR = 0.5 (1 + r.x/|r|)  -- the same
// continuous function the reference evaluates through two atan2f and three cosf (its float 90-degree constant differs
// from pi/2 by 4e-8, i.e. 2e-8 in a gain); no truncation follows the pan gains, so nothing here needs the reference's
// bits.  A source exactly at the listener (or straight above / below) gives 0.5 / 0.5 like the reference.  This is what
// the CUDA kernel uses (K1 is bound by the length of its dependent instruction chain: 0.103 -> 0.095 ms per million
// sources); the host emulation keeps the reference's sequence of operations below.
SBM_HD void stereo_pan_gains_direct(const float* r, float* gl, float* gr) {
	const float n2 = fadd(fadd(fmul(r[0], r[0]), fmul(r[1], r[1])), fmul(r[2], r[2]));
	const float s = n2 > 0.0f ? fdiv(r[0], fsqrt(n2)) : 0.0f;
	*gl = fmul(0.5f, fsub(1.0f, s));
	*gr = fmul(0.5f, fadd(1.0f, s));
}

// StereoPanner gains (RA/dsp/stereo_panner.cc:28-44).  cosf is continuous here (no truncation follows), so the
// device's cosf (<= 2 ulp) is as good as libm's.
SBM_HD void stereo_pan_gains(float az, float el, float* gl, float* gr) {
	const float kRadFromDeg = (float)(3.14159265358979323846 / 180.0);
	const float left = fmul(kRadFromDeg, 90.0f), right = fmul(kRadFromDeg, -90.0f);
	const float ce = cosf(el);
	*gl = fmul(0.5f, fadd(1.0f, fmul(cosf(fsub(left, az)), ce)));
	*gr = fmul(0.5f, fadd(1.0f, fmul(cosf(fsub(right, az)), ce)));
}

// AmbisonicLookupTable::GetEncodingCoeffs, order 3, spread 0 (RA/ambisonics/ambisonic_lookup_table.cc:88-161).
// `table` = 91*91*15 floats built on the host (sh_table.cc).  Symmetry signs (:200-216) in closed form:
//   left/right flip (azimuth < 0): -1 for m < 0;   up/down (elevation < 0): (-1)^(n+m);
//   front/back (|azimuth| > 90): m < 0 ? -(-1)^|m| : (-1)^m.
SBM_HD void encoding_coeffs(const float* table, float az, float el, float* out16) {
	const float kPi = (float)3.14159265358979323846;
	const float kDegFromRad = (float)(180.0 / 3.14159265358979323846);
	const int az_deg = az < kPi ? (int)fmul(az, kDegFromRad) : (int)fmul(az, kDegFromRad) - 360;
	const int el_deg = (int)fmul(el, kDegFromRad);
	const int abs_az = az_deg < 0 ? -az_deg : az_deg;
	const int az_idx = abs_az > 90 ? 180 - abs_az : abs_az;
	const int el_idx = el_deg < 0 ? -el_deg : el_deg;
	const float* cell = table + (az_idx * kTableEl + el_idx) * kShRaw;
	out16[0] = 1.0f;
	int raw = 0;
	for (int n = 1; n <= 3; ++n)
		for (int m = -n; m <= n; ++m, ++raw) {
			const int am = m < 0 ? -m : m;
			float flip = 1.0f;
			if (az_deg < 0 && m < 0) flip = -1.0f;
			if (el_deg < 0 && ((n + m) & 1)) flip = -flip;
			if (abs_az > 90) { const float fb = (am & 1) ? -1.0f : 1.0f; flip = fmul(flip, m < 0 ? -fb : fb); }
			out16[raw + 1] = fmul(cell[raw], flip);
		}
}

// Converts a non-negative coordinate to 32.32 fixed point (floor).
SBM_HD unsigned long long to_fx(double x) { return (unsigned long long)(x * kFx); }

// Largest window (in samples) the single-mix-source fast form may ask for; longer reads (source_delta >~ 4 at 1024 frames)
// stay on the generic per-sample path.
constexpr uint32_t kMaxMix1Window = 4096;

// Returns kUsed* bits; *rendered_src = PCM samples a rendered source reads this buffer (0 when skipped / stopped).
// `rec_out`, `shg_out`, `shr_out`: where this source's WorkRec / ShGains / ShRamps go (the arrays in HBM indexed by render-list
// position, or -- fused kernel -- registers and the CTA's shared-memory rings).
// `mix_hint`: SrcStatic::mix_first of this slot as the render list knows it (SlotArrays::list_mix), or 0xffffffff.
SBM_HD uint32_t update_source(const Globals& G, const SlotArrays& A, DevFlags* flags, const float* table, uint32_t slot, uint32_t mix_hint, WorkRec* rec_out,
                              ShGains* shg_out, ShRamps* shr_out, ReflSend* rfl_out, uint32_t* rendered_src) {
	*rendered_src = 0;
	// (the first mix source's parameters and read head, requested together with the slot's state)
	MixZ mz_hint; double pos_hint = 0.0;
	mz_hint.pcm = nullptr; mz_hint.len = 0; mz_hint.delta = 0.0; mz_hint.mix_factor = 0.0f; mz_hint.pad = 0.0;
	if (mix_hint != 0xffffffffu) { mz_hint = A.mixz[mix_hint]; pos_hint = A.mix_pos[mix_hint]; }
	// All per-slot state is loaded up front, unconditionally: the loads are independent of each other, so the memory
	// system sees them as one batch instead of a chain of dependent round trips behind the branches below.
	const SrcStatic st = A.stat[slot];
	const uint32_t live = A.live[slot];
	const F4 pv = A.posvol[slot];
	const float occ_t = A.occl[slot];
	const unsigned long long read_i0 = A.read_i[slot];
	long long eof = A.eof[slot];
	Resamp rs = A.rs[slot];
	float dg = A.dgain[slot];
	float occ = A.occ_cur[slot];
	const float lpf_prev0 = A.lpf_prev[slot];
	uint8_t bits = A.bits[slot];
	const int prev_used_sh = flags->sh_used[G.parity ^ 1], prev_used_pan = flags->pan_used[G.parity ^ 1];   // previous buffer's mixers (GainMixer sweep)
	const int prev_used_refl = G.room ? flags->refl_used[G.parity ^ 1] : 0;
	float rg = G.room ? A.rgain[slot] : 0.0f;
	// stereo-pan bus gains (the first two of the slot's cstride floats), loaded with the rest: left where they are used,
	// the compiler must keep them behind the state stores above them (possible aliasing) = one more DRAM round trip
	// (quad-major layout: channels 4q .. 4q+3 of slot s are the F4 at cgain4[q * cstride + s], so a warp's slots are consecutive
	// 16-byte elements; the pan gains are the first two floats of quad 0)
	F4* cgain4 = reinterpret_cast<F4*>(A.cgain);
	const F4 cg_q0 = cgain4[slot];
	const float cg_pan0 = cg_q0.x, cg_pan1 = cg_q0.y;
	WorkRec rec;
	rec.pcm = st.pcm; rec.len = st.len; rec.start = 0; rec.n_src = 0; rec.avail = 0; rec.flags = kRecSkip;
	rec.slot = slot; rec.x0_fx = 0; rec.step_fx = 0; rec.prev0 = 0.0f; rec.prev1 = 0.0f; rec.lpf_coef = 0.0f; rec.lpf_prev = 0.0f;
	rec.ch_ramp_mask = 0; rec.mix_factor = 1.0f; rec.mix_first = st.mix_first; rec.nz = st.nz;
	rec.direct.konst = 0.0f; rec.direct.start = 0.0f; rec.direct.inc = 0.0f; rec.direct.n = 0;
	rec.ch[0] = rec.direct; rec.ch[1] = rec.direct;
	if (!(st.flags & kFlagAlive) || !live) { *rec_out = rec; return 0; }

	const uint32_t F = G.frames;
	const uint32_t kind = (st.flags >> kFlagKindShift) & 0xff;
	const uint32_t feed = (st.flags >> kFlagFeedShift) & 3;
	const bool looping = (st.flags & kFlagLooping) != 0;
	uint32_t rflags = 0;

	// ---------------- feeder: how many source samples, from where (AE.cpp:378-474) ----------------
	uint32_t n_src = F;
	uint32_t alg_samples = 0;       // PCM samples this buffer has to read (accounting only)
	const bool resample = st.rate != G.rate;
	if (resample) {
		// AudioResampler::numSrcSamplesNeeded (AR.cpp:77-86)
		const int64_t max_dest = rs.prev_dest_dst_coords + (int64_t)F;
		const double max_dest_src = (double)max_dest * (double)st.rate / (double)G.rate;
		const int64_t prev1 = rs.prev_samples_0_src_coords + 1;
		n_src = (uint32_t)(int)((uint64_t)ceil(max_dest_src) - (uint64_t)prev1);
		// AudioResampler::resample coordinates (AR.cpp:101-103)
		const int64_t dest0 = rs.prev_dest_dst_coords + 1;
		const double dest0_src = (double)dest0 * (double)st.rate / (double)G.rate;
		const double x0 = dest0_src - (double)rs.prev_samples_0_src_coords;
		rec.x0_fx = to_fx(x0 > 0.0 ? x0 : 0.0);
		rec.step_fx = to_fx((double)st.rate / (double)G.rate);
		rec.prev0 = rs.prev_samples[0]; rec.prev1 = rs.prev_samples[1];
		rs.prev_dest_dst_coords += (int64_t)F;
		rs.prev_samples_0_src_coords += (int64_t)n_src;
		// prev_samples[] for the next buffer are the last two samples of {prev0, prev1, window}: written by K2.
		A.rs[slot].prev_dest_dst_coords = rs.prev_dest_dst_coords;
		A.rs[slot].prev_samples_0_src_coords = rs.prev_samples_0_src_coords;
		rflags |= kRecResample;
	}
	bool remove_source = false;
	rec.n_src = n_src;
	if (feed == kFeedBuffer) {
		// AE.cpp:389-426
		uint64_t cur = read_i0;
		const uint64_t len = st.len;
		rec.start = (uint32_t)(cur < len ? cur : len);
		if (cur + n_src <= len) {
			rec.avail = n_src;
			cur += n_src;
			if (looping && cur == len) cur = 0;
		} else if (cur < len) {
			if (looping) { rec.avail = n_src; rflags |= kRecWrap; cur = (cur + n_src) % len; }
			else { rec.avail = (uint32_t)(len - cur); cur = len; }
			remove_source = (st.flags & kFlagRemoveOnFinish) && cur >= len;
		} else {
			rec.avail = 0;                                   // read index parked at/after the end: zeros
			remove_source = (st.flags & kFlagRemoveOnFinish) != 0;
		}
		A.read_i[slot] = cur;
		alg_samples = rec.avail;
	} else if (feed == kFeedRing) {
		// AE.cpp:459-474: pop what is there, zero-pad the rest.  read_i counts samples popped so far.
		const uint64_t rd = read_i0;
		const uint64_t wr = A.ring_write[slot];
		const uint64_t have = wr - rd < (uint64_t)n_src ? wr - rd : (uint64_t)n_src;
		rec.start = (uint32_t)(rd % st.len);
		rec.avail = (uint32_t)have;
		rflags |= kRecWrap;
		A.read_i[slot] = rd + have;
		alg_samples = rec.avail;
	} else {
		// AE.cpp:435-458: every mix source's read head advances by source_delta before each of the n_src reads.
		bool fast = st.nz == 1;
		for (uint32_t z = 0; z < st.nz; ++z) {
			const uint32_t k = st.mix_first + z;
			const bool hinted = z == 0 && mix_hint == st.mix_first;
			const MixZ mz = hinted ? mz_hint : A.mixz[k];
			const double p0 = hinted ? pos_hint : A.mix_pos[k];
			A.mix_pos0[k] = p0;
			// The reference adds delta n_src times in double; one fma differs by < n_src ulps of pos (continuous).
			const double pend = fma((double)n_src, mz.delta, p0);
			A.mix_pos[k] = pend;
			alg_samples += (uint32_t)((double)n_src * mz.delta) + 2u;
			if (fast) {
				// Single mix source as a contiguous window: first read at p1 = p0 + delta, last at pend; window sample j is
				// pcm[(floor(p1) + j) % len], the read head starts at frac(p1) inside it and steps by delta.
				const double p1 = p0 + mz.delta;
				const double span = pend - floor(p1);
				if (!(mz.delta > 0.0) || !(p1 >= 0.0) || !(span < (double)kMaxMix1Window) || !(pend < 9.0e15)) { fast = false; continue; }
				const unsigned long long base = (unsigned long long)p1;
				const uint32_t need = (uint32_t)span + 3u;              // +1: the interpolation partner, +1: fixed-point slack
				if (need > mz.len) { fast = false; continue; }
				rec.pcm = mz.pcm; rec.len = mz.len;
				// 32-bit remainder whenever the read head is below 2^32 samples (a day of audio): 64-bit division is ~5x the cost
				rec.start = base <= 0xffffffffull ? (uint32_t)base % mz.len : (uint32_t)(base % mz.len);
				rec.n_src = need; rec.avail = need;
				rec.x0_fx = to_fx(p1 - (double)base);
				rec.step_fx = to_fx(mz.delta);
				rec.mix_factor = mz.mix_factor;
				if (rec.start + need > mz.len) rflags |= kRecWrap;
			}
		}
		rflags |= fast ? kRecMix1 : kRecMix;
		if (!fast) { rec.n_src = n_src; rec.avail = 0; rflags &= ~kRecWrap; }
	}

	// ---------------- EOF bookkeeping (AE.cpp:496-517) ----------------
	eof -= (long long)n_src;
	A.eof[slot] = eof;
	if (remove_source || (!looping && eof <= 0)) {
		// DestroySource runs from the task queue before the graph is processed: the last buffer is not rendered.
		A.live[slot] = 0;
		A.finished[finished_push(flags)] = (int)st.id;
		*rec_out = rec;
		return 0;
	}

	// ---------------- attenuation + GainNode (distance_attenuation.cc:82-119, gain_node.cc:38-80) ----------------
	const float src_pos[3] = {pv.x, pv.y, pv.z};
	const float input_gain = fmul(G.master, pv.w);
	const bool hrtf = kind == kKindHrtf;
	uint32_t used_refl = 0;
	if (G.room && kind != kKindStereo) {
		// The reflections send (graph_manager.cc:150, gain_mixer_node.cc:44-63): every fed sound object adds its RAW buffer to the
		// mono reflections mixer through a ramped gain = room effects gain (1) * direct attenuation * ReflectionProperties::gain
		// (distance_attenuation.cc:112-113), whatever the GainNode decides below.  Its GainMixer sweeps every buffer it had input.
		if (prev_used_refl) {
			if (bits & kBitReflActive) bits &= (uint8_t)~kBitReflActive;
			else bits &= (uint8_t)~kBitReflExists;
		}
		const float direct_att = fmul(input_gain, log_distance_attenuation(G.head_pos, src_pos));
		const float refl_target = fmul(fmul(1.0f, direct_att), G.refl_gain);
		const RampDesc d = plan_gain(&rg, (bits & kBitReflExists) != 0, refl_target, F, /*accumulate=*/true);
		bits |= kBitReflExists | kBitReflActive;
		A.rgain[slot] = rg;
		rfl_out->konst = d.konst; rfl_out->start = d.start; rfl_out->inc = d.inc; rfl_out->n = d.n;
		used_refl = kUsedRefl;
	}
	if (kind != kKindStereo) {
		// GainMixer::Reset sweep (gain_mixer.cc:80-96): it runs once between two non-empty buffers of that mixer and
		// touches every source that owns processors, whether or not the source reaches the mixer this buffer.
		const int prev_used = hrtf ? prev_used_sh : prev_used_pan;
		if (prev_used) {
			if (bits & kBitMixActive) bits &= (uint8_t)~kBitMixActive;
			else bits &= (uint8_t)~kBitMixExists;
		}
	}
	float target;
	if (kind == kKindStereo) target = input_gain;                                        // AttenuationType::kInput
	else target = fmul(input_gain, log_distance_attenuation(G.head_pos, src_pos));        // kDirect
	if (near_zero(target) && near_zero(dg)) {
		A.dgain[slot] = 0.0f; A.bits[slot] = bits | kBitDirectInit;
		// skip: nothing downstream runs, but the feeder still resampled this buffer (AE.cpp:485-494 precede the graph), so
		// K2 must still carry AudioResampler::prev_samples forward.
		rec.flags = kRecSkip | (rflags & (kRecResample | kRecWrap));
		*rec_out = rec;
		return used_refl;     // (its reflections send is below -60 dB too: K2 adds nothing for a skipped source, the mixer still counts as fed)
	}
	if (near_unity(target) && near_unity(dg)) {
		dg = 1.0f;
		rec.direct.konst = 1.0f; rec.direct.start = 1.0f; rec.direct.inc = 0.0f; rec.direct.n = 0;
	} else {
		rec.direct = plan_gain(&dg, (bits & kBitDirectInit) != 0, target, F, /*accumulate=*/false);
	}
	bits |= kBitDirectInit;
	A.dgain[slot] = dg;
	*rendered_src = alg_samples ? alg_samples : 1u;

	if (kind == kKindStereo) {
		rflags |= kRecStereo;
		rec.ch[0].konst = 1.0f; rec.ch[1].konst = 1.0f;
		A.bits[slot] = bits;
		rec.flags = rflags;
		*rec_out = rec;
		return kUsedStereo;
	}

	// ---------------- OcclusionNode (occlusion_node.cc:53-99) ----------------
	occ = fadd(occ_t, fmul(kOcclusionSmoothing, fsub(occ, occ_t)));
	A.occ_cur[slot] = occ;
	rec.lpf_coef = occlusion_coefficient(occ);
	rec.lpf_prev = lpf_prev0;
	if (!(rec.lpf_coef < kEpsilonFloat)) rflags |= kRecLpf;

	// ---------------- direction + mixer gains ----------------
	float az = 0.0f, el = 0.0f;
#if defined(__CUDA_ARCH__)
	float rel[3];
	if (hrtf) source_direction(G.head_pos, G.head_q, src_pos, &az, &el);
	else source_relative(G.head_pos, G.head_q, src_pos, rel);
#else
	source_direction(G.head_pos, G.head_q, src_pos, &az, &el);
#endif
	const bool chan_init = (bits & kBitMixExists) != 0;     // processors exist => already initialised by an ApplyGain
	bits |= kBitMixExists | kBitMixActive;
	A.bits[slot] = bits;
	if (!hrtf) {
		float g[2];
#if defined(__CUDA_ARCH__)
		stereo_pan_gains_direct(rel, &g[0], &g[1]);
#else
		stereo_pan_gains(az, el, &g[0], &g[1]);
#endif
		float curs[2];
		for (int c = 0; c < 2; ++c) {
			float cur = c == 0 ? cg_pan0 : cg_pan1;
			rec.ch[c] = plan_gain(&cur, chan_init, g[c], F, /*accumulate=*/true);
			curs[c] = cur;
			if (rec.ch[c].n) rec.ch_ramp_mask |= 1u << c;
		}
		cgain4[slot] = F4{curs[0], curs[1], cg_q0.z, cg_q0.w};
	} else {
		float g[kNumSh];
		encoding_coeffs(table, az, el, g);
		// Four channels at a time: the slot's 16 bus gains are read and written back as four 16-byte accesses, and the
		// effective constant gains leave as four 16-byte stores (read and written one float at a time, the 16 loads were
		// 16 dependent round trips -- each sits behind the previous iteration's stores, which may alias as far as the compiler
		// knows -- and the 4-byte stores of a warp scatter over 32 sectors each).
		F4* sg4 = reinterpret_cast<F4*>(shg_out->konst);
		for (int q = 0; q < kNumSh / 4; ++q) {
			F4* cgq = cgain4 + (size_t)q * (size_t)A.cstride + slot;
			const F4 v = q == 0 ? cg_q0 : *cgq;
			float cur[4] = {v.x, v.y, v.z, v.w}, kon[4], stm[4], rin[4], rnf[4];
			uint32_t gm = 0;
			for (int j = 0; j < 4; ++j) {
				const int c = 4 * q + j;
				const RampDesc d = plan_gain(&cur[j], chan_init, g[c], F, /*accumulate=*/true);
				kon[j] = d.konst; stm[j] = 0.0f; rin[j] = 0.0f; rnf[j] = 0.0f;
				if (d.n) {
					stm[j] = fsub(d.start, d.konst); rin[j] = d.inc; rnf[j] = (float)d.n;
					gm |= 1u << j;
					if (d.n > rec.ch[0].n) rec.ch[0].n = d.n;       // ambisonic sources: ch[0].n = longest channel ramp (K2's ramp region)
					// ended on a near-zero value that the -60 dB rule zeroes: konst is 0, not the ramp's end value.  (An end value
					// below 1e-6 -- a coefficient whose target is exactly 0, reached to within rounding -- is as good as 0.)
					if (d.n < F && d.konst != cur[j] && fabsf(cur[j]) > 1e-6f) rflags |= kRecRampCareful;
				}
			}
			*cgq = F4{cur[0], cur[1], cur[2], cur[3]};
			sg4[q] = F4{kon[0], kon[1], kon[2], kon[3]};
			if (gm) {
				// ramp parameters of the group, 16 bytes each (zeros for the channels of the group that do not ramp)
				reinterpret_cast<F4*>(shr_out->stm)[q] = F4{stm[0], stm[1], stm[2], stm[3]};
				reinterpret_cast<F4*>(shr_out->inc)[q] = F4{rin[0], rin[1], rin[2], rin[3]};
				reinterpret_cast<F4*>(shr_out->nf)[q] = F4{rnf[0], rnf[1], rnf[2], rnf[3]};
				rec.ch_ramp_mask |= gm << (4 * q);
			}
		}
		// K2 reads all 48 ramp parameters of a source that has any channel ramp: the groups without one must say 0
		if (rec.ch_ramp_mask)
			for (int q = 0; q < kNumSh / 4; ++q)
				if (!((rec.ch_ramp_mask >> (4 * q)) & 0xfu)) {
					const F4 z = F4{0.0f, 0.0f, 0.0f, 0.0f};
					reinterpret_cast<F4*>(shr_out->stm)[q] = z; reinterpret_cast<F4*>(shr_out->inc)[q] = z; reinterpret_cast<F4*>(shr_out->nf)[q] = z;
				}
		rflags |= kRecHrtf;
	}
	rec.flags = rflags;
	*rec_out = rec;
	return (hrtf ? kUsedSh : kUsedPan) | used_refl;
}

}  // namespace sbm
