// k2_rec.h -- from a source's WorkRec (k1_core.h: what the feeder and Resonance's per-source parameter logic decided for this
// buffer) to its MixRec (sbm_core.h: what the mix kernel's lean paths consume, in final form): how the window is laid out in
// a stage slot of the TMA ring, which copies fetch it, which code path renders it, pre-multiplied constant gains.
//
// Evaluated ONCE per source and buffer by whoever produces the records -- the K1 kernel (k1_params.cu) or the parameter warps
// of the fused kernel (k2_mix.cu) -- so that the warp that drives the TMA ring has nothing left to decide: it reads the
// 16-byte fetch plan and issues the copies.  (Round 1 planned windows and condensed records in that warp; ncu showed the math
// warps waiting for it -- a serial, latency-bound chain of ~300 instructions per batch -- not for HBM.)
#pragma once
#include "k2_core.h"

namespace sbm {

// Stage slot per source, in floats: room for a window of frames * 1.25 (sample-rate ratios <= 1, doppler / pitch up
// to +25 %) plus alignment slack; larger windows take the careful path.
SBM_HD uint32_t window_slot(uint32_t F) { return (F + F / 4 + 24u + 3u) & ~3u; }

// ---------------------------------------------------------------------------------------------------------------
// How a source's window is laid out in its stage slot.  Pure function of the record.
//   slot: [4 floats unused][TMA span ...]; span = pcm[a0 .. a1) with a0 = start & ~3.  Window sample j sits at
//   w[j] for j < n1 and at w[j + wadj] after a loop wrap (second segment placed at the next 16-byte boundary).
// ---------------------------------------------------------------------------------------------------------------
struct WinPlan {
	bool fast;           // window comes through the TMA ring
	uint32_t off0;       // start & 3
	uint32_t seg1_floats, seg2_floats;   // 16-byte multiples; seg2 = 0 without wrap
	uint32_t a0;         // first fetched sample of segment 1
	uint32_t n1;         // window samples before the wrap (0xffffffff: no wrap)
	uint32_t wadj;       // index adjustment after the wrap
};

SBM_HD WinPlan plan_window(const WorkRec& r, uint32_t wslot) {
	WinPlan p;
	p.fast = false; p.off0 = r.start & 3u; p.a0 = r.start & ~3u; p.seg1_floats = 0; p.seg2_floats = 0; p.n1 = 0xffffffffu; p.wadj = 0;
	const uint32_t flags = r.flags;
	if (flags & (kRecSkip | kRecMix)) return p;
	if (r.avail != r.n_src || r.n_src == 0 || r.n_src > r.len) return p;
	if ((reinterpret_cast<unsigned long long>(r.pcm) & 15ull) != 0) return p;
	const uint32_t end = r.start + r.n_src;
	if (end <= r.len) {
		p.seg1_floats = ((end + 3u) & ~3u) - p.a0;
	} else {
		if (!(flags & kRecWrap)) return p;
		const uint32_t len4 = (r.len + 3u) & ~3u;
		p.n1 = r.len - r.start;
		p.wadj = len4 - r.len;
		p.seg1_floats = len4 - p.a0;
		p.seg2_floats = (r.n_src - p.n1 + 3u) & ~3u;
	}
	if (p.seg1_floats + p.seg2_floats + 4u > wslot) return p;
	p.fast = true;
	return p;
}

// The careful paths (and a gain-skipped resampled source, whose history must still be carried forward) read the full WorkRec
// back from global memory; every other source is rendered from its MixRec alone, and its WorkRec need not be written at all.
SBM_HD bool mixrec_needs_workrec(uint32_t mode) {
	return (mode & (kMdGeneric | kMdLpf | kMdAdjust)) != 0 || ((mode & kMdSkip) && (mode & kMdResample));
}

SBM_HD MixRec make_mixrec(const WorkRec& r, uint32_t wslot, bool sh) {
	const WinPlan p = plan_window(r, wslot);
	MixRec m;
	const uint32_t f = r.flags;
	const bool rs = (f & kRecResample) != 0;
	// the careful path evaluates every case with its own tests: windows that are not staged, re-aligned loop wraps, the
	// low-pass, and channel ramps that end on a value the -60 dB rule zeroes
	const bool careful = !p.fast || (f & kRecRampCareful) != 0;
	m.mode = 0;
	if (f & kRecSkip) m.mode |= kMdSkip;
	if (careful) m.mode |= kMdGeneric;
	if (f & (kRecResample | kRecMix1)) m.mode |= kMdInterp;
	if (rs) m.mode |= kMdResample;
	if (p.wadj) m.mode |= kMdAdjust;
	if (f & kRecLpf) m.mode |= kMdLpf;
	if (f & kRecStereo) m.mode |= kMdStereo;
	m.w_off = 4u + p.off0 - (rs ? 2u : 0u);
	m.n1b = p.n1 == 0xffffffffu ? 0xffffffffu : p.n1 + (rs ? 2u : 0u);
	m.wadj = p.wadj;
	m.x0 = r.x0_fx; m.step = r.step_fx;
	// mix_factor: folded into the constant gains on the lean paths; the careful paths apply it to the samples themselves
	const float mf = (!careful && (f & kRecMix1) && !p.wadj && !(f & kRecLpf)) ? r.mix_factor : 1.0f;
	m.gd = mf * r.direct.konst;
	m.g0 = m.gd * r.ch[0].konst; m.g1 = m.gd * r.ch[1].konst;
	uint32_t nr = r.direct.n;
	if (!sh) {
		if ((r.ch_ramp_mask & 1u) && r.ch[0].n > nr) nr = r.ch[0].n;
		if ((r.ch_ramp_mask & 2u) && r.ch[1].n > nr) nr = r.ch[1].n;
	} else if (r.ch_ramp_mask) {
		m.mode |= kMdShRamp;
		if (r.ch[0].n > nr) nr = r.ch[0].n;      // ambisonic sources: K1 leaves the longest channel ramp in ch[0].n
	}
	m.nramp = nr;
	if (nr) m.mode |= kMdRamp;
	m.slot = r.slot;
	m.mfx = (f & kRecMix1) ? r.mix_factor : 1.0f;
	m.prev0 = r.prev0; m.prev1 = r.prev1;
	m.d_konst = r.direct.konst; m.d_start = r.direct.start; m.d_inc = r.direct.inc;
	m.dn_nsrc = (r.direct.n & 0xffffu) | ((r.n_src & 0xffffu) << 16);
	m.c0_konst = r.ch[0].konst; m.c0_start = r.ch[0].start; m.c0_inc = r.ch[0].inc;
	m.c1_konst = r.ch[1].konst; m.c1_start = r.ch[1].start; m.c1_inc = r.ch[1].inc;
	// (stereo-bus sources only: an ambisonic source's ch[0].n carries its longest channel ramp, not a ramp of ch[0])
	m.cn = sh ? 0u : (((r.ch_ramp_mask & 1u) ? (r.ch[0].n & 0xffffu) : 0u) | (((r.ch_ramp_mask & 2u) ? (r.ch[1].n & 0xffffu) : 0u) << 16));
	m.pad = 0;
	m.pcm = r.pcm; m.a0 = p.a0;
	m.segs = p.fast ? (p.seg1_floats | (p.seg2_floats << 16)) : 0u;
	return m;
}

// Ramp evaluation from a MixRec (GainProcessor::ApplyGain reduced to a descriptor, sbm_core.h plan_gain / k2_core.h gain_at).
SBM_HD float direct_gain_at(const MixRec& m, uint32_t i) { return i < (m.dn_nsrc & 0xffffu) ? fmaf((float)i, m.d_inc, m.d_start) : m.d_konst; }
SBM_HD float ch0_gain_at(const MixRec& m, uint32_t i) { return i < (m.cn & 0xffffu) ? fmaf((float)i, m.c0_inc, m.c0_start) : m.c0_konst; }
SBM_HD float ch1_gain_at(const MixRec& m, uint32_t i) { return i < (m.cn >> 16) ? fmaf((float)i, m.c1_inc, m.c1_start) : m.c1_konst; }

}  // namespace sbm
