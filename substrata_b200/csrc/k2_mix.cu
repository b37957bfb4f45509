// k2_mix.cu -- "K2": the hot kernel.  For every rendered source of this buffer: gather its sample window from HBM
// (loop / ring wrap, linear sample-rate conversion, pitch / doppler read heads), apply the ramped direct gain and the
// optional occlusion low-pass, multiply by the source's 2 (stereo pan / non-spatial) or 16 (3rd-order ambisonic)
// ramped bus gains and accumulate into the output tile; and "K3", the fixed-order cross-CTA reduction of the partial
// buses into this rank's bus.
//
// Decomposition.  A CTA owns ALL frames of the buffer and every grid-th batch of the render list, and keeps its partial
// sums in registers across the whole slice, so per-source contributions never leave the register file; each CTA then
// stores one partial bus [NCH][frames] and K3 sums the partial buses in CTA order (deterministic; no atomics on audio).
//   stereo bus (2 channels):  one WARP per source -- lane l owns frames l, l+32, ... as (L, R) accumulator pairs (one FFMA2
//                             per frame; the interpolations of two frames share packed f32x2 instructions); the four math
//                             warps of a CTA take the eight sources of a batch in two rounds and combine their tiles
//                             through shared memory once, at the end of the slice;
//   ambisonic bus (16 ch):    one CTA per batch of four sources -- thread t owns frames t, t+T, t+2T, t+3T and all 16
//                             channels (32 channel-pair accumulators, FFMA2 with the gain pair from an LDS.128 and the
//                             sample broadcast); a batch of simple sources runs as one straight-line block.
//
// Data movement.  Source windows are not loaded by the math threads: a dedicated ISSUE warp drives a multi-stage
// shared-memory ring with the TMA unit -- one cp.async.bulk (UBLKCP) per window segment, completion signalled on an
// mbarrier per stage, stages handed back through a second mbarrier by the math warps.  The per-source records arrive in
// FINAL form (MixRec, 128 B, sbm_core.h / k2_rec.h: render path, read head, pre-multiplied gains, ramps, and the 16-byte
// fetch plan -- pointer, first sample, segment lengths) through a second TMA ring two batches deeper; the issue warp's lane
// b reads the plan of source b and issues its copies, nothing else.  (Round 1 planned the windows and condensed the records
// in that warp; ncu showed the math warps waiting on its serial ~300-instruction chain per batch, not on HBM.)
// Windows start at arbitrary sample offsets; the copy fetches the enclosing 16-byte-aligned span and the math indexes
// from the sub-offset.  A looping window that crosses the end of its buffer is fetched as two segments.
// Sources the lean paths cannot express (zero-padded ends, ring underflow, multi-source pitch mixes, windows larger
// than a stage slot, unaligned buffers, occlusion low-pass) take a careful per-sample path that reads the full WorkRec and,
// where the window is not staged, global memory.
//
// HBM traffic: 4 B per source sample (PCM, + <= 24 B alignment over-fetch per segment) + 128 B MixRec (+64 B SH gains)
// per source; writes are O(grid * frames).  Roofline: HBM read.  No tensor cores: 2..16 FMA per 4 B loaded.
#include <map>
#include <mutex>

#include "device_internal.h"
#include "k2_rec.h"

namespace sbm {

namespace {

#ifdef SBM_CONTIGUOUS   // development A/B: contiguous slice per CTA instead of round-robin batches
#define SBM_GLOBAL_BATCH(jb) (blockIdx.x * ((((s1 - s0) + kBatch - 1) / kBatch + gridDim.x - 1) / gridDim.x) + (jb))
#else
#define SBM_GLOBAL_BATCH(jb) (blockIdx.x + (jb) * gridDim.x)
#endif
#ifndef SBM_STAGES
#define SBM_STAGES 2
#endif
#ifndef SBM_MIN_CTAS
#define SBM_MIN_CTAS 4
#endif
#ifndef SBM_PACKED_LERP
#define SBM_PACKED_LERP 1             // stereo kernel: interpolate frame pairs with packed f32x2 instructions
#endif
#ifndef SBM_PRODUCER_SLEEP_NS
#define SBM_PRODUCER_SLEEP_NS 200
#endif
#ifndef SBM_SH_FUSE
#define SBM_SH_FUSE 2
#endif
#ifndef SBM_SH_BIG_FUSE
#define SBM_SH_BIG_FUSE 2
#endif
#ifndef SBM_SH_BIG_CTAS
#define SBM_SH_BIG_CTAS 1             // resident CTAs per SM the 1024-frame ambisonic kernel's register budget is set for (2 -> 96 registers and spills)
#endif
#ifndef SBM_BATCH_PAN
#define SBM_BATCH_PAN 8
#endif
// Sources per pipeline stage.  The producer warp plans one source per LANE, so its (latency-bound, serial) cost per batch
// is the same for 4 or 8 sources: the stereo kernel, whose math warps take a source each, uses batches of 8 (two rounds
// of its four warps) to halve the producer's cost per source -- it was the producer, not HBM, the math warps waited for.
#ifndef SBM_BATCH_SH
#define SBM_BATCH_SH 4
#endif
#ifndef SBM_BATCH_SH_BIG
#define SBM_BATCH_SH_BIG 8            // sources per stage of the 1024-frame ambisonic shape.  It is alone on its SM (one issue warp, eight math warps): the per-batch
#endif                                // costs -- the issue warp's serial chain, the math warps' barrier round trip and record scan -- are amortised over twice the
                                      // sources: 0.150 -> 0.137 ms per 65 536 sources (0.108 -> 0.084 with the ramp fix-ups compiled out).  The 512-frame shape
                                      // (three CTAs per SM) measures the same with 4 and 8
constexpr int kBatchSh = SBM_BATCH_SH, kBatchPan = SBM_BATCH_PAN, kPanWarps = 4;
// Window stages in flight per CTA (batches of kBatch sources); the record ring is two batches deeper.
#ifndef SBM_STAGES_SH
#define SBM_STAGES_SH 4
#endif
#ifndef SBM_STAGES_SH_BIG
#define SBM_STAGES_SH_BIG 6
#endif
constexpr int kStagesPan = SBM_STAGES, kStagesSh = SBM_STAGES_SH, kStagesShBig = SBM_STAGES_SH_BIG;
constexpr int kMaxStages = 10;

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + TMA bulk copy (global -> shared), sm_90+ / sm_100a.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
	uint32_t ok;
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
	    "selp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(ok)
	    : "r"(smem_u32(bar)), "r"(parity)
	    : "memory");
	return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
	uint32_t spins = 0;
	while (!mbar_try_wait(bar, parity)) {
		if (++spins > (1u << 24)) __trap();    // a lost completion must fail loudly, not hang the device
	}
}
// Producer-side wait: the producer runs several batches ahead of the math warps, so it spends most of its life waiting
// for a stage to be handed back.  Sleeping between polls keeps it out of the issue slots of the math warps that share
// its scheduler (a hot spin cost ~70 issue slots per source on that scheduler).
__device__ __forceinline__ void mbar_wait_relaxed(unsigned long long* bar, uint32_t parity) {
	uint32_t spins = 0;
	while (!mbar_try_wait(bar, parity)) {
		__nanosleep(SBM_PRODUCER_SLEEP_NS);
		if (++spins > (1u << 22)) __trap();
	}
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
	             "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Packed pair of fp32 FMAs (Blackwell FFMA2): acc.xy += g.xy * x.xy in one issue slot.
__device__ __forceinline__ void ffma2(float2& acc, const float2 g, const float2 x) {
	unsigned long long a = *reinterpret_cast<unsigned long long*>(&acc);
	asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(*reinterpret_cast<const unsigned long long*>(&g)), "l"(*reinterpret_cast<const unsigned long long*>(&x)));
	acc = *reinterpret_cast<float2*>(&a);
}

__device__ __forceinline__ float2 fadd2(const float2 a, const float2 b) {
	unsigned long long r;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
	return *reinterpret_cast<float2*>(&r);
}

struct MixSmemHdr {
	unsigned long long full_w[kMaxStages], empty_w[kMaxStages];          // window ring: TMA landed / math warps done
	unsigned long long full_r[kMaxStages + 2], empty_r[kMaxStages + 2];  // record ring
	float lpf_carry[8];                                          // per math warp (stereo kernel) / [0] (ambisonic kernel)
	unsigned int done[kMaxStages + 2];                           // self-service rings: warps that have finished the batch in record slot q
};
static_assert(sizeof(MixSmemHdr) <= 512, "header must fit its 512-byte slot");

// Barrier among the math warps only (the producer warp never joins): named barrier 1.
__device__ __forceinline__ void math_sync(uint32_t n_threads) { asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory"); }
__device__ __forceinline__ void math_sync_id(uint32_t id, uint32_t n_threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory"); }

// Element d of {prev0, prev1, window...} from the staged window, with the reference's index clamp (AR.cpp:68-71).
__device__ __forceinline__ float staged_tmp(const WorkRec& r, const float* w, const WinPlan& p, int d) {
	const int max_i = (int)r.n_src + 1;
	d = d > max_i ? max_i : d;
	if (d == 0) return r.prev0;
	if (d == 1) return r.prev1;
	const uint32_t j = (uint32_t)(d - 2);
	return w[j + (j >= p.n1 ? p.wadj : 0u)];
}

// Engine-rate input sample i from the staged window, every case handled (careful path; w = window sample 0).
__device__ __forceinline__ float staged_sample(const WorkRec& r, const float* w, const WinPlan& p, uint32_t i) {
	if (!(r.flags & (kRecResample | kRecMix1))) return w[i + (i >= p.n1 ? p.wadj : 0u)];
	int d; float fr;
	fx_split(r.x0_fx, r.step_fx, i, &d, &fr);
	float v0, v1;
	if (r.flags & kRecResample) { v0 = staged_tmp(r, w, p, d); v1 = staged_tmp(r, w, p, d + 1); }
	else {
		const uint32_t j0 = (uint32_t)d, j1 = (uint32_t)d + 1u;
		v0 = w[j0 + (j0 >= p.n1 ? p.wadj : 0u)]; v1 = w[j1 + (j1 >= p.n1 ? p.wadj : 0u)];
		return fmaf(fr, v1 - v0, v0) * r.mix_factor;
	}
	return fmaf(fr, v1 - v0, v0);
}

// Same from a MixRec on the lean resample path: wb = element 0 of {prev0, prev1, window...} in the stage slot (wb[0], wb[1]
// themselves are not staged: they come from the record).
__device__ __forceinline__ float staged_tmp_m(const MixRec& m, const float* wb, int d) {
	const int max_i = (int)(m.dn_nsrc >> 16) + 1;
	d = d > max_i ? max_i : d;
	if (d == 0) return m.prev0;
	if (d == 1) return m.prev1;
	return wb[(uint32_t)d + ((uint32_t)d >= m.n1b ? m.wadj : 0u)];
}

// Occlusion low-pass over `F` samples staged in shared memory, by ONE warp.  y[n] = c (y[n-1] - x[n]) + x[n] is linear
// in y[n-1]: the warp splits the buffer into 32 contiguous chunks, runs each chunk from a zero state while accumulating
// c^len, chains the 32 chunk states in order, and re-runs each chunk from its true initial state.  Returns the final state.
__device__ float lpf_warp(float* buf, uint32_t F, float coef, float prev, uint32_t lane) {
	const uint32_t chunk = (F + 31) / 32;
	const uint32_t lo = lane * chunk < F ? lane * chunk : F;
	const uint32_t hi = lo + chunk < F ? lo + chunk : F;
	float p = 0.0f, apow = 1.0f;
	for (uint32_t n = lo; n < hi; ++n) { p = lpf_step(coef, p, buf[n]); apow *= coef; }
	float carry = prev, mine = prev;
	for (int l = 0; l < 32; ++l) {
		const float pl = __shfl_sync(0xffffffffu, p, l), al = __shfl_sync(0xffffffffu, apow, l);
		if ((int)lane == l) mine = carry;
		carry = fmaf(al, carry, pl);
	}
	float q = mine;
	for (uint32_t n = lo; n < hi; ++n) { q = lpf_step(coef, q, buf[n]); buf[n] = q; }
	return carry;
}

// Shared-memory carve-up (both kernels): header | MixRec ring | ShGains ring, ShRamps ring (ambisonic only) | window ring.
template <int NCH, int kStages, int kBatch>
struct SmemView {
	static constexpr int kRecStages = kStages + 2;
	MixSmemHdr* hdr; MixRec* recs; ReflSend* rfls; ShGains* shgs; ShRamps* shrs; float* wring;
	__device__ explicit SmemView(unsigned char* raw, bool room) {
		hdr = reinterpret_cast<MixSmemHdr*>(raw);
		recs = reinterpret_cast<MixRec*>(raw + 512);
		unsigned char* p = reinterpret_cast<unsigned char*>(recs) + sizeof(MixRec) * kRecStages * kBatch;
		rfls = reinterpret_cast<ReflSend*>(p);      // room effects: the reflections send ramps, staged with the records
		if (room) p += sizeof(ReflSend) * kRecStages * kBatch;
		shgs = reinterpret_cast<ShGains*>(p);
		if (NCH == 16) p += sizeof(ShGains) * kRecStages * kBatch;
		shrs = reinterpret_cast<ShRamps*>(p);       // channel ramps of the sources that have any, staged with the windows
		if (NCH == 16) p += sizeof(ShRamps) * kStages * kBatch;
		wring = reinterpret_cast<float*>(p);
	}
};

// Overlapped K1 / K2 (k1_params.cu): before the records of render-list entries [b0, b0 + nb) are fetched, wait until K1 has stamped their
// tiles.  One lane polls (acquire, gpu scope); the bulk copies that follow read through the async proxy, hence the proxy fence.  K1 is
// resident and never waits for this kernel, so the wait is bounded by K1's own progress; the spin limit (~2 s) only turns a lost K1
// (a fault) into an error status instead of a hung stream.
__device__ __forceinline__ void wait_records_ready(const MixArgs& a, uint32_t b0, uint32_t nb) {
	const uint32_t t0 = b0 / kK1Tile, t1 = (b0 + nb - 1) / kK1Tile;
	for (uint32_t t = t0; t <= t1; ++t) {
		uint32_t spins = 0;
		for (;;) {
			uint32_t v;
			asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.ready + t) : "memory");
			if (v == a.epoch) break;
			__nanosleep(64);
			++spins;
			if ((spins & 1023u) == 0 && *reinterpret_cast<const volatile int*>(&a.flags->pipe_timeout)) break;      // somebody gave up already: drain
			if (spins > (1u << 24)) { const_cast<DevFlags*>(a.flags)->pipe_timeout = 1; __threadfence(); break; }
		}
	}
	asm volatile("fence.proxy.async.global;" ::: "memory");
}

// ================= issue warp (both kernels): lane b looks after source b of each batch =================
template <int NCH, int kStages, int kBatch>
__device__ void issue_warp(const SmemView<NCH, kStages, kBatch>& sm, const MixArgs& a, uint32_t s0, uint32_t s1, uint32_t nbatch, uint32_t wslot, uint32_t lane) {
	constexpr int kRecStages = kStages + 2;
	MixSmemHdr* hdr = sm.hdr;
	const MixRec* __restrict__ grec = a.mrec;
	const ShGains* __restrict__ gshg = a.shg;
	// Records run two batches ahead of the windows.  Per batch: (1) wait for its records and read the fetch plans, (2) refill the
	// record slot two batches ahead, (3) wait for the window stage to be handed back, (4) arm the barrier and issue the copies.
	auto issue_recs = [&](uint32_t jb, uint32_t q, uint32_t parity_prev) {
		if (jb >= nbatch) return;
		if (jb >= (uint32_t)kRecStages) mbar_wait_relaxed(&hdr->empty_r[q], parity_prev);
		if (lane == 0) {
			const uint32_t b0 = s0 + SBM_GLOBAL_BATCH(jb) * kBatch;
			const uint32_t nb = s1 - b0 < (uint32_t)kBatch ? s1 - b0 : (uint32_t)kBatch;
			const uint32_t bytes = nb * (uint32_t)sizeof(MixRec) + (NCH == 16 ? nb * (uint32_t)sizeof(ShGains) : 0u) + (a.rfl ? nb * (uint32_t)sizeof(ReflSend) : 0u);
			if (a.ready) wait_records_ready(a, b0, nb);
			mbar_expect_tx(&hdr->full_r[q], bytes);
			bulk_g2s(sm.recs + q * kBatch, grec + b0, nb * (uint32_t)sizeof(MixRec), &hdr->full_r[q]);
			if (NCH == 16) bulk_g2s(sm.shgs + q * kBatch, gshg + b0, nb * (uint32_t)sizeof(ShGains), &hdr->full_r[q]);
			if (a.rfl) bulk_g2s(sm.rfls + q * kBatch, a.rfl + b0, nb * (uint32_t)sizeof(ReflSend), &hdr->full_r[q]);
		}
		__syncwarp();
	};
	issue_recs(0, 0, 0);
	issue_recs(1, 1, 0);
	uint32_t qr = 0, ph_r = 0;        // record slot of batch jb, its full_r parity
	uint32_t qw = 0, ph_w = 0;        // window stage of batch jb, its full_w parity
	uint32_t q2 = 2, ph_2 = 0;        // record slot of batch jb + 2, its full_r parity
	for (uint32_t jb = 0; jb < nbatch; ++jb) {
		mbar_wait(&hdr->full_r[qr], ph_r);
		const uint32_t b0 = s0 + SBM_GLOBAL_BATCH(jb) * kBatch;
		const uint32_t nb = s1 - b0 < (uint32_t)kBatch ? s1 - b0 : (uint32_t)kBatch;
		uint4 pl = make_uint4(0u, 0u, 0u, 0u);     // {pcm lo, pcm hi, a0, segs}
		bool want_ramps = false;
		if (lane < nb) {
			const MixRec& m = sm.recs[qr * kBatch + lane];
			pl = *reinterpret_cast<const uint4*>(&m.pcm);
			if (NCH == 16) want_ramps = (m.mode & (kMdShRamp | kMdSkip)) == kMdShRamp;
		}
		const uint32_t seg1 = pl.w & 0xffffu, seg2 = pl.w >> 16;
		uint32_t bytes = 4u * (seg1 + seg2);
		for (int o = kBatch / 2; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
		// channel ramps: ONE copy for the batch's ShRamps when any of its sources has some (the entries of sources without ramps are
		// stale and never read) -- a copy per source costs this warp ~12 serial instructions each
		const bool any_ramps = NCH == 16 && __ballot_sync(0xffffffffu, want_ramps) != 0u;
		if (any_ramps) bytes += nb * (uint32_t)sizeof(ShRamps);
		issue_recs(jb + 2, q2, ph_2 ^ 1u);
		if (jb >= (uint32_t)kStages) mbar_wait_relaxed(&hdr->empty_w[qw], ph_w ^ 1u);
		__syncwarp();
		if (lane == 0) {
			if (bytes) mbar_expect_tx(&hdr->full_w[qw], bytes); else mbar_arrive(&hdr->full_w[qw]);
			if (any_ramps) bulk_g2s(sm.shrs + qw * kBatch, a.shr + b0, nb * (uint32_t)sizeof(ShRamps), &hdr->full_w[qw]);
		}
		__syncwarp();
		if (seg1) {
			const float* pcm = reinterpret_cast<const float*>((unsigned long long)pl.x | ((unsigned long long)pl.y << 32));
			float* dst = sm.wring + (size_t)(qw * kBatch + lane) * wslot + 4;
			bulk_g2s(dst, pcm + pl.z, 4u * seg1, &hdr->full_w[qw]);
			if (seg2) bulk_g2s(dst + seg1, pcm, 4u * seg2, &hdr->full_w[qw]);
		}
		__syncwarp();
		if (++qr == (uint32_t)kRecStages) { qr = 0; ph_r ^= 1u; }
		if (++qw == (uint32_t)kStages) { qw = 0; ph_w ^= 1u; }
		if (++q2 == (uint32_t)kRecStages) { q2 = 0; ph_2 ^= 1u; }
	}
}

// ================= self-service rings (ambisonic kernel, frames 256 / 512 / 1024): no dedicated issue warp =================
// With the records in final form, refilling a stage is ~40 instructions -- too little to justify a warp that holds registers
// the math warps need (the register file is per scheduler: 4 x 16 K; a 5th or 9th warp per CTA costs a whole resident CTA at
// 128 registers).  Every warp is a math warp; the warp that finishes a batch LAST (elected through a shared-memory counter)
// refills the window stage and the record slot that batch just vacated.  Nobody ever waits for a stage to be handed back.
template <int NCH, int kStages, int kBatch>
__device__ __forceinline__ void fill_record_slot(const SmemView<NCH, kStages, kBatch>& sm, const MixArgs& a, uint32_t s0, uint32_t s1, uint32_t nbatch, uint32_t lane, uint32_t j) {
	constexpr int kRecStages = kStages + 2;
	if (j >= nbatch) return;
	if (lane == 0) {
		const uint32_t q = j % kRecStages;
		const uint32_t b0 = s0 + SBM_GLOBAL_BATCH(j) * kBatch;
		const uint32_t nb = s1 - b0 < (uint32_t)kBatch ? s1 - b0 : (uint32_t)kBatch;
		const uint32_t bytes = nb * (uint32_t)sizeof(MixRec) + (NCH == 16 ? nb * (uint32_t)sizeof(ShGains) : 0u) + (a.rfl ? nb * (uint32_t)sizeof(ReflSend) : 0u);
		mbar_expect_tx(&sm.hdr->full_r[q], bytes);
		bulk_g2s(sm.recs + q * kBatch, a.mrec + b0, nb * (uint32_t)sizeof(MixRec), &sm.hdr->full_r[q]);
		if (NCH == 16) bulk_g2s(sm.shgs + q * kBatch, a.shg + b0, nb * (uint32_t)sizeof(ShGains), &sm.hdr->full_r[q]);
		if (a.rfl) bulk_g2s(sm.rfls + q * kBatch, a.rfl + b0, nb * (uint32_t)sizeof(ReflSend), &sm.hdr->full_r[q]);
	}
	__syncwarp();
}
template <int NCH, int kStages, int kBatch>
__device__ __forceinline__ void fill_window_stage(const SmemView<NCH, kStages, kBatch>& sm, const MixArgs& a, uint32_t s0, uint32_t s1, uint32_t nbatch, uint32_t wslot, uint32_t lane, uint32_t j) {
	constexpr int kRecStages = kStages + 2;
	if (j >= nbatch) return;
	MixSmemHdr* hdr = sm.hdr;
	const uint32_t qr = j % kRecStages, qw = j % kStages;
	mbar_wait(&hdr->full_r[qr], (j / kRecStages) & 1u);     // the batch's records (requested RS - S batches ago)
	const uint32_t b0 = s0 + SBM_GLOBAL_BATCH(j) * kBatch;
	const uint32_t nb = s1 - b0 < (uint32_t)kBatch ? s1 - b0 : (uint32_t)kBatch;
	uint4 pl = make_uint4(0u, 0u, 0u, 0u);     // {pcm lo, pcm hi, a0, segs}
	bool want_ramps = false;
	if (lane < nb) {
		const MixRec& m = sm.recs[qr * kBatch + lane];
		pl = *reinterpret_cast<const uint4*>(&m.pcm);
		if (NCH == 16) want_ramps = (m.mode & (kMdShRamp | kMdSkip)) == kMdShRamp;
	}
	const uint32_t seg1 = pl.w & 0xffffu, seg2 = pl.w >> 16;
	uint32_t bytes = 4u * (seg1 + seg2) + (want_ramps ? (uint32_t)sizeof(ShRamps) : 0u);
	for (int o = kBatch / 2; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
	if (lane == 0) {
		if (bytes) mbar_expect_tx(&hdr->full_w[qw], bytes); else mbar_arrive(&hdr->full_w[qw]);
	}
	__syncwarp();
	if (seg1) {
		const float* pcm = reinterpret_cast<const float*>((unsigned long long)pl.x | ((unsigned long long)pl.y << 32));
		float* dst = sm.wring + (size_t)(qw * kBatch + lane) * wslot + 4;
		bulk_g2s(dst, pcm + pl.z, 4u * seg1, &hdr->full_w[qw]);
		if (seg2) bulk_g2s(dst + seg1, pcm, 4u * seg2, &hdr->full_w[qw]);
	}
	if (NCH == 16 && want_ramps) bulk_g2s(sm.shrs + qw * kBatch + lane, a.shr + (b0 + lane), (uint32_t)sizeof(ShRamps), &hdr->full_w[qw]);
	__syncwarp();
}

template <int kStages>
__device__ __forceinline__ void init_barriers(MixSmemHdr* hdr, uint32_t n_consumers) {
	constexpr int kRecStages = kStages + 2;
	for (int q = 0; q < kStages; ++q) { mbar_init(&hdr->full_w[q], 1); mbar_init(&hdr->empty_w[q], n_consumers); }
	for (int q = 0; q < kRecStages; ++q) { mbar_init(&hdr->full_r[q], 1); mbar_init(&hdr->empty_r[q], n_consumers); hdr->done[q] = 0u; }
	fence_mbar_init();
}

// Stereo kernel, every case the lean forms cannot express (zero-padded ends, ring underflow, multi-source pitch mixes, oversize
// windows, unaligned buffers, re-aligned loop wraps, occlusion low-pass): OUT OF LINE, one warp per source, straightforward
// arithmetic.  Leaves the source's engine-rate samples in its own stage slot -- slot[i], i < frames, mix_factor applied; the
// direct gain and the low-pass too when it returns true -- and does the feeder / filter bookkeeping; the caller then reads the
// slot like a native-rate window.  Keeping this (large, rarely executed) code out of the kernel body shrinks the hot path's
// instruction footprint; its staging array lives in local memory, which only this path touches.
// apply_lpf = false (room effects): the occlusion low-pass is left to direct_gain_and_lpf_in_slot, so that the caller can take
// the RAW samples for the reflections send first.
__device__ __noinline__ bool careful_source_to_slot(const MixArgs& a, const WorkRec& r, float* slot, uint32_t wslot, uint32_t lane, bool apply_lpf) {
	const uint32_t F = a.frames;
	const WinPlan wp = plan_window(r, wslot);
	const float* w = slot + 4 + wp.off0;
	float x[kMaxFrames / 32];     // held until every lane has read the staged window: the result overwrites it
#pragma unroll 1
	for (uint32_t k = 0; k * 32u < F; ++k) {
		const uint32_t i = lane + 32u * k;
		x[k] = 0.0f;
		if (i < F) x[k] = wp.fast ? staged_sample(r, w, wp, i) : source_sample(r, a.mixz, a.mix_pos0, i);
	}
	if ((r.flags & kRecResample) && lane == 0) {
		float p0, p1;
		if (wp.fast) { p0 = staged_tmp(r, w, wp, (int)r.n_src); p1 = staged_tmp(r, w, wp, (int)r.n_src + 1); }
		else resampler_tail(r, &p0, &p1);
		a.rs[r.slot].prev_samples[0] = p0; a.rs[r.slot].prev_samples[1] = p1;
	}
	__syncwarp();
	const bool lpf = apply_lpf && (r.flags & kRecLpf) != 0;
#pragma unroll 1
	for (uint32_t k = 0; k * 32u < F; ++k) {
		const uint32_t i = lane + 32u * k;
		if (i < F) slot[i] = lpf ? x[k] * gain_at(r.direct, i) : x[k];
	}
	__syncwarp();
	if (lpf) {
		// direct gain applied above; the recurrence runs in place in the stage slot
		const float carry = lpf_warp(slot, F, r.lpf_coef, r.lpf_prev, lane);
		__syncwarp();
		if (lane == 0) a.lpf_prev[r.slot] = carry;
	}
	fence_proxy_async();     // generic-proxy writes to the slot precede the TMA refill
	return lpf;
}

// Second half of the careful path for a low-passed source when the first half ran with apply_lpf = false: slot[i] holds the raw
// samples; multiply by the direct gain and run the occlusion low-pass in place.
__device__ __noinline__ void direct_gain_and_lpf_in_slot(const MixArgs& a, const WorkRec& r, float* slot, uint32_t lane) {
	const uint32_t F = a.frames;
	for (uint32_t i = lane; i < F; i += 32u) slot[i] = slot[i] * gain_at(r.direct, i);
	__syncwarp();
	const float carry = lpf_warp(slot, F, r.lpf_coef, r.lpf_prev, lane);
	__syncwarp();
	if (lane == 0) a.lpf_prev[r.slot] = carry;
	fence_proxy_async();
}

template <int V> struct IntTag { static constexpr int value = V; };
__device__ __forceinline__ float refl_gain_at(const ReflSend& s, uint32_t i) { return i < s.n ? fmaf((float)i, s.inc, s.start) : s.konst; }

__device__ __forceinline__ float frac_of(unsigned long long p) { return __uint_as_float(0x3f800000u | ((uint32_t)p >> 9)) - 1.0f; }

// =====================================================================================================================
// Stereo-bus kernel (panned + non-spatial sources): kBatch math warps, one source per warp, + the producer warp.
// KF = frames / 32 rounded up to 8, 16 or 32; EXACT = (frames == 32 * KF): no per-frame bounds checks.
// =====================================================================================================================
// ROOM (room effects on): a third accumulator per frame collects the source's RAW signal through the ramped reflections send
// gain (graph_manager.cc:150); the CTA's partial bus then has three rows.
// STAGES: window stages in flight per CTA.  kStagesPan with SBM_MIN_CTAS resident CTAs per SM is the stand-alone shape; the overlapped
// shape (K1 resident beside this kernel, k1_params.cu) runs one CTA fewer per SM with one stage more -- the same bytes in flight -- which
// leaves K1's CTAs their registers (the register file is per scheduler: 4 x 16 K).
constexpr int kStagesPanOverlap = kStagesPan + 1;
template <int KF, bool EXACT, bool ROOM, int STAGES>
__global__ void __launch_bounds__(kPanWarps * 32 + 32, KF == 32 ? 2 : (ROOM ? 3 : SBM_MIN_CTAS)) k2_mix_stereo_kernel(MixArgs a, uint32_t first, uint32_t n, uint32_t wslot, float* __restrict__ partials) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	constexpr int kStages = STAGES, kRecStages = kStages + 2, kBatch = kBatchPan;
	const SmemView<2, kStages, kBatch> sm(smem_raw, ROOM);
	MixSmemHdr* hdr = sm.hdr;
	// EXACT: the frame count is a compile-time constant (every "is this frame F - 1 / inside the buffer" test folds away)
	const uint32_t F = EXACT ? 32u * (uint32_t)KF : a.frames, t = threadIdx.x, lane = t & 31u, warp = t >> 5;
	// Batches of kBatch consecutive list entries are dealt to the CTAs round-robin (batch g -> CTA g % grid), so that a
	// run of expensive sources (long ramps, occlusion, generic fetches) is spread over the whole grid.
	const uint32_t s0 = first, s1 = first + n;
	const uint32_t total_batches = (n + kBatch - 1) / kBatch;
#ifdef SBM_CONTIGUOUS
	const uint32_t per_cta = (total_batches + gridDim.x - 1) / gridDim.x;
	const uint32_t nbatch = blockIdx.x * per_cta < total_batches ? (total_batches - blockIdx.x * per_cta < per_cta ? total_batches - blockIdx.x * per_cta : per_cta) : 0u;
#else
	const uint32_t nbatch = blockIdx.x < total_batches ? (total_batches - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
#endif
	if (t == 0) init_barriers<kStages>(hdr, kPanWarps);
	__syncthreads();
	// everything above overlaps the tail of K1; its records are read from here on -- tile by tile as K1 stamps them when it runs beside us
	if (!a.ready) grid_dep_wait();
	if (warp == (uint32_t)kPanWarps) { issue_warp<2, kStages, kBatch>(sm, a, s0, s1, nbatch, wslot, lane); return; }

	float2 acc[KF];     // (L, R) of frame lane + 32 k: one FFMA2 (gain pair x broadcast sample) per frame on the constant-gain path
	float acc3[ROOM ? KF : 1];     // the reflections send
#pragma unroll
	for (int k = 0; k < KF; ++k) acc[k] = make_float2(0.0f, 0.0f);
#pragma unroll
	for (int k = 0; k < (ROOM ? KF : 1); ++k) acc3[k] = 0.0f;

	uint32_t qr = 0, qw = 0, ph_w = 0;     // record slot / window stage of batch jb, parity of its full_w
	for (uint32_t jb = 0; jb < nbatch; ++jb, qr = qr + 1 == (uint32_t)kRecStages ? 0u : qr + 1, ph_w ^= (qw + 1 == (uint32_t)kStages), qw = qw + 1 == (uint32_t)kStages ? 0u : qw + 1) {
		mbar_wait(&hdr->full_w[qw], ph_w);
		const uint32_t b0 = s0 + SBM_GLOBAL_BATCH(jb) * kBatch;
#pragma unroll 1
		for (uint32_t src = warp; src < (uint32_t)kBatch; src += kPanWarps)     // this warp's sources of the batch
		if (b0 + src < s1) {
			const MixRec& m = sm.recs[qr * kBatch + src];
			const uint32_t mode = m.mode;
			float* slot = sm.wring + (size_t)(qw * kBatch + src) * wslot;      // this source's stage slot
			if (mode & kMdSkip) {
				// gain-skipped source: the feeder still consumed this buffer; carry the resampler history forward
				if ((mode & kMdResample) && lane == 0) {
					const WorkRec& r = a.rec[b0 + src];
					float p0, p1;
					resampler_tail(r, &p0, &p1);
					a.rs[r.slot].prev_samples[0] = p0; a.rs[r.slot].prev_samples[1] = p1;
				}
			} else {
				float x[KF];           // engine-rate input samples of this lane's frames, before the direct gain
				float mf = 1.0f;       // MixSource::mix_factor where it is not already inside x
				bool filtered = false; // x already carries the direct gain and the occlusion low-pass
				if (EXACT && SBM_PACKED_LERP && (mode & ~kMdRamp) == kMdInterp) {
					// ---------------- single read head (pitch / doppler), window staged by TMA, nothing else: the C4 case.  All
					// frames in pairs, the two interpolations of a pair share packed f32x2 instructions. ----------------
					const float* wb = slot + m.w_off;
					mf = m.mfx;
					const unsigned long long step = m.step;
					unsigned long long p = m.x0 + (unsigned long long)lane * step;
					const unsigned long long step32 = step << 5;
#pragma unroll
					for (int k = 0; k < KF; k += 2) {
						const unsigned long long pa = p, pb = p + step32;
						const uint32_t da = (uint32_t)(pa >> 32), db = (uint32_t)(pb >> 32);
						const float2 v0 = make_float2(wb[da], wb[db]), v1 = make_float2(wb[da + 1], wb[db + 1]);
						const float2 fb = make_float2(__uint_as_float(0x3f800000u | ((uint32_t)pa >> 9)), __uint_as_float(0x3f800000u | ((uint32_t)pb >> 9)));
						const float2 fr = fadd2(fb, make_float2(-1.0f, -1.0f));
						float2 dv = v1;
						ffma2(dv, v0, make_float2(-1.0f, -1.0f));     // v1 - v0
						float2 v = v0;
						ffma2(v, fr, dv);                              // v0 + fr (v1 - v0)
						x[k] = v.x; x[k + 1] = v.y;
						p += step32 << 1;
					}
				} else if (mode & (kMdGeneric | kMdLpf | kMdAdjust)) {
					// ---------------- careful fetch (rare, out of line): leaves the samples in this source's stage slot ----------------
					filtered = careful_source_to_slot(a, a.rec[b0 + src], slot, wslot, lane, !ROOM);
#pragma unroll
					for (int k = 0; k < KF; ++k) { const uint32_t i = lane + 32u * k; x[k] = (EXACT || i < F) ? slot[i] : 0.0f; }
					if (ROOM && (mode & kMdLpf)) {
						// room effects: the reflections send takes the raw samples; then direct gain + low-pass, and read again
						const ReflSend rs = sm.rfls[qr * kBatch + src];
#pragma unroll
						for (int k = 0; k < KF; ++k) acc3[k] = fmaf(refl_gain_at(rs, lane + 32u * k), x[k], acc3[k]);
						__syncwarp();
						direct_gain_and_lpf_in_slot(a, a.rec[b0 + src], slot, lane);
#pragma unroll
						for (int k = 0; k < KF; ++k) { const uint32_t i = lane + 32u * k; x[k] = (EXACT || i < F) ? slot[i] : 0.0f; }
						filtered = true;
					}
				} else {
					// ---------------- lean fetch: window staged in shared memory by the TMA ring ----------------
					const float* wb = slot + m.w_off;
					mf = m.mfx;
					if (mode & kMdInterp) {
						const bool rsmp = (mode & kMdResample) != 0;
						const unsigned long long step = m.step;
						unsigned long long p = m.x0 + (unsigned long long)lane * step;
						const unsigned long long step32 = step << 5;
						auto lerp1 = [&](int k) {
							const uint32_t i = lane + 32u * k;
							float v = 0.0f;
							if (EXACT || i < F) {
								const uint32_t di = (uint32_t)(p >> 32);
								const float fr = frac_of(p);
								float v0, v1;
								// only the first frames (resampler history) and the last ones (index clamp) need the careful form
								if ((k == 0 || k == KF - 1 || !EXACT) && rsmp) { v0 = staged_tmp_m(m, wb, (int)di); v1 = staged_tmp_m(m, wb, (int)di + 1); }
								else { v0 = wb[di]; v1 = wb[di + 1]; }
								v = fmaf(fr, v1 - v0, v0);
							}
							x[k] = v;
							p += step32;
						};
						if (EXACT && SBM_PACKED_LERP) {
							// frames 1 .. KF-2 in pairs: the two interpolations of a pair share packed f32x2 instructions
							lerp1(0);
#pragma unroll
							for (int k = 1; k + 1 < KF - 1; k += 2) {
								const unsigned long long pa = p, pb = p + step32;
								const uint32_t da = (uint32_t)(pa >> 32), db = (uint32_t)(pb >> 32);
								const float2 v0 = make_float2(wb[da], wb[db]), v1 = make_float2(wb[da + 1], wb[db + 1]);
								const float2 fb = make_float2(__uint_as_float(0x3f800000u | ((uint32_t)pa >> 9)), __uint_as_float(0x3f800000u | ((uint32_t)pb >> 9)));
								const float2 fr = fadd2(fb, make_float2(-1.0f, -1.0f));
								float2 dv = v1;
								ffma2(dv, v0, make_float2(-1.0f, -1.0f));     // v1 - v0
								float2 v = v0;
								ffma2(v, fr, dv);                              // v0 + fr (v1 - v0)
								x[k] = v.x; x[k + 1] = v.y;
								p += step32 << 1;
							}
							lerp1(KF - 1);
						} else {
#pragma unroll
							for (int k = 0; k < KF; ++k) lerp1(k);
						}
						if (rsmp && lane == 0) {
							const int ns = (int)(m.dn_nsrc >> 16);
							a.rs[m.slot].prev_samples[0] = staged_tmp_m(m, wb, ns);
							a.rs[m.slot].prev_samples[1] = staged_tmp_m(m, wb, ns + 1);
						}
					} else {
#pragma unroll
						for (int k = 0; k < KF; ++k) {
							const uint32_t i = lane + 32u * k;
							x[k] = (EXACT || i < F) ? wb[i] : 0.0f;
						}
					}
				}
				// ---------------- direct gain + bus gains (GainNode, GainMixer in accumulate mode) ----------------
				const float2 g01 = make_float2(m.g0, m.g1);
				const uint32_t nramp = filtered ? F : m.nramp;     // frames that need per-frame gain evaluation
				const bool refl_here = ROOM && !(filtered && (mode & kMdLpf)) && !(mode & kMdStereo);     // (a low-passed source sent its raw samples above)
				ReflSend rs;
				rs.konst = 0.0f; rs.start = 0.0f; rs.inc = 0.0f; rs.n = 0;
				if (refl_here) rs = sm.rfls[qr * kBatch + src];
				const float gr = rs.konst * mf;     // constant part of the send gain, on the sample as x holds it
				// RAMPS: 0 = no gain ramps in this buffer, 1 = ramps end within the first 32 frames (the usual case while the
				// listener moves), 2 = anything (long ramps, filtered signal): per-iteration warp-uniform test.
				auto gains = [&](auto ramps_tag) {
					constexpr int RAMPS = decltype(ramps_tag)::value;
#pragma unroll
					for (int k = 0; k < KF; ++k) {
						const uint32_t i = lane + 32u * k;
						const bool per_frame = RAMPS == 2 ? (32u * k < nramp) : (RAMPS == 1 && k == 0);
						if (per_frame) {
							const float yv = filtered ? x[k] : (x[k] * mf) * direct_gain_at(m, i);
							acc[k].x = fmaf(ch0_gain_at(m, i), yv, acc[k].x);
							acc[k].y = fmaf(ch1_gain_at(m, i), yv, acc[k].y);
							if (!filtered && !(mode & kMdStereo) && i == F - 1) a.lpf_prev[m.slot] = yv;
							if (ROOM && refl_here) acc3[k] = fmaf(refl_gain_at(rs, i), x[k] * mf, acc3[k]);
						} else {
							ffma2(acc[k], g01, make_float2(x[k], x[k]));
							if (ROOM && refl_here) acc3[k] = fmaf(gr, x[k], acc3[k]);
							// bypassed occlusion filter remembers the last input sample (mono_pole_filter.cc:40-44)
							if ((k == KF - 1 || !EXACT) && !(mode & kMdStereo) && i == F - 1) a.lpf_prev[m.slot] = x[k] * m.gd;
						}
					}
				};
				if (nramp == 0) gains(IntTag<0>());
				else if (nramp <= 32u) gains(IntTag<1>());
				else gains(IntTag<2>());
			}
		}
		// this warp is done with window stage qw and record slot qr: hand them back to the producer
		__syncwarp();
		if (lane == 0) { mbar_arrive(&hdr->empty_w[qw]); mbar_arrive(&hdr->empty_r[qr]); }
	}
	// ---- combine the math warps' tiles (the window ring is idle now: every batch has been consumed) ----
	math_sync(kPanWarps * 32);
	constexpr uint32_t ROWS = ROOM ? 3u : 2u;     // L, R (+ the reflections send)
	float* tile = sm.wring;     // [kPanWarps][ROWS][F]
#pragma unroll
	for (int k = 0; k < KF; ++k) {
		const uint32_t i = lane + 32u * k;
		if (EXACT || i < F) {
			tile[(warp * ROWS + 0) * F + i] = acc[k].x; tile[(warp * ROWS + 1) * F + i] = acc[k].y;
			if (ROOM) tile[(warp * ROWS + 2) * F + i] = acc3[k];
		}
	}
	math_sync(kPanWarps * 32);
	float* out = partials + (size_t)blockIdx.x * ROWS * F;
	for (uint32_t q = t; q < ROWS * F; q += kPanWarps * 32) {
		float s = 0.0f;
#pragma unroll
		for (int wv = 0; wv < kPanWarps; ++wv) s += tile[wv * ROWS * F + q];
		out[q] = s;
	}
}

// =====================================================================================================================
// Ambisonic-bus kernel: T math threads (frames / 4) work on one batch of sources at a time + the producer warp.
//   Thread t owns frames t, t + T, t + 2T, t + 3T and all 16 bus channels of them: 64 accumulators as 32 channel PAIRS,
//   one FFMA2 (gain pair x broadcast sample) per pair and frame; the gains come out of shared memory as LDS.128 = two
//   ready-made pairs.  A batch whose four sources are all "simple" (native-rate or single read head, no filter, ramps
//   inside the first T frames) runs as ONE straight-line block -- descriptor and sample loads of all four sources
//   first, then the arithmetic -- because the per-source dependent chain (load descriptor -> branch -> load samples
//   -> FMA) is what bounds this kernel, not the issue rate.  The frames inside a gain ramp are the first few of the
//   buffer = iteration k = 0 of the first warps: those warps evaluate the 16 ramping gains in registers, the others
//   never see the ramp.
// =====================================================================================================================
#ifndef SBM_SH_BIG_K
#define SBM_SH_BIG_K 4                // frames per thread of the > 512-frame ambisonic kernel (2 = 16 math warps per CTA measured 4 % slower: per-warp overhead scales with the warp count)
#endif
#ifndef SBM_SH_BIG_MAXNREG
#define SBM_SH_BIG_MAXNREG 96         // register cap of the two-group shape (see SBM_SH_BIG_G)
#endif
#ifndef SBM_SH_BIG_G
#define SBM_SH_BIG_G 1                // math groups per CTA of the > 512-frame ambisonic kernel.  2 (two groups sharing the producer, each taking every
                                      // other batch) needs 17 warps x >= 112 registers, which the SM's allocation granularity (CTAs rounded up to
                                      // 20 warps) refuses: "too many resources requested for launch"; at the 96 registers that do fit it spills
#endif
constexpr int kShK = 4, kShBigK = SBM_SH_BIG_K, kShBigT = 1024 / kShBigK, kShBigG = SBM_SH_BIG_G;
// G: independent groups of TMAX math threads in one CTA.  Each group has its own accumulators and takes every G-th batch of
// the CTA's slice; the groups share the producer warp and its rings.  Where the register file leaves room for only one CTA
// per SM (frames > 512: 256 math threads x ~120 registers), two groups overlap each other's barrier waits and dependent
// chains the way two resident CTAs would.
template <int TMAX, int K, int G, bool EXACT, bool SELF, int kStages, bool ROOM, int kBatch = kBatchSh>     // TMAX math threads x K frames each; EXACT: frames == K * TMAX,
                                                                          // no per-frame bounds checks; SELF: self-service rings (no issue warp; implies EXACT, G == 1);
                                                                          // ROOM: a 17th accumulator row = the reflections send (room effects); kBatch: sources per stage
__device__ __forceinline__ void ambisonic_body(const MixArgs& a, uint32_t first, uint32_t n, uint32_t wslot, float* __restrict__ partials) {
	constexpr int NCH = 16;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	constexpr int kRecStages = kStages + 2;
	const SmemView<16, kStages, kBatch> sm(smem_raw, ROOM);
	MixSmemHdr* hdr = sm.hdr;
	// blockDim.x = T math threads (+ one issue warp, the last warp, unless SELF)
	const uint32_t F = a.frames, T = EXACT ? (uint32_t)TMAX : (blockDim.x - 32) / G;   // EXACT: T is a compile-time constant
	const uint32_t grp = G == 1 ? 0u : (threadIdx.x / T < (uint32_t)G ? threadIdx.x / T : 0u);      // math group of this thread
	const uint32_t lane = threadIdx.x & 31u;
	// Which 32-frame chunk a warp owns is rotated from CTA to CTA (SELF): gain ramps live in the first frames of a buffer, so
	// the warp that owns chunk 0 does the fix-up work of every ramping source and is the slowest of its CTA; warp w of every
	// CTA sits on scheduler w % 4, and without the rotation the slow warps of all resident CTAs would share one scheduler
	// while the other three idle (ncu: 26 % of the stall samples were warps of the light chunks waiting for a stage).
	const uint32_t rot = SELF ? (blockIdx.x / a.rot_div) % (T / 32u) : 0u;
	const uint32_t t = SELF ? (((threadIdx.x >> 5) + rot) % (T / 32u)) * 32u + lane : threadIdx.x - grp * T;     // index inside the group = frame of iteration k = 0
	// (making F a compile-time constant too was measured 5 % slower here: it tips the register allocation into spills)
	// Batches of kBatch consecutive list entries are dealt to the CTAs round-robin (batch g -> CTA g % grid), so that a
	// run of expensive sources (long ramps, occlusion, generic fetches) is spread over the whole grid.
	const uint32_t s0 = first, s1 = first + n;
	const uint32_t total_batches = (n + kBatch - 1) / kBatch;
#ifdef SBM_CONTIGUOUS
	const uint32_t per_cta = (total_batches + gridDim.x - 1) / gridDim.x;
	const uint32_t nbatch = blockIdx.x * per_cta < total_batches ? (total_batches - blockIdx.x * per_cta < per_cta ? total_batches - blockIdx.x * per_cta : per_cta) : 0u;
#else
	const uint32_t nbatch = blockIdx.x < total_batches ? (total_batches - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
#endif
	if (threadIdx.x == 0) init_barriers<kStages>(hdr, T / 32);     // a batch is consumed by the T / 32 warps of ONE group
	__syncthreads();
	grid_dep_wait();     // everything above overlaps the tail of the previous kernel; K1's records are read from here on
	if (!SELF && threadIdx.x >= (uint32_t)G * T) { issue_warp<16, kStages, kBatch>(sm, a, s0, s1, nbatch, wslot, lane); return; }
	if (SELF && threadIdx.x < 32) {     // (hardware warp 0, whatever chunk it owns)
		// prologue of the self-service rings: the first record slots and window stages
		for (uint32_t j = 0; j < (uint32_t)kRecStages; ++j) fill_record_slot<16, kStages, kBatch>(sm, a, s0, s1, nbatch, lane, j);
		for (uint32_t j = 0; j < (uint32_t)kStages; ++j) fill_window_stage<16, kStages, kBatch>(sm, a, s0, s1, nbatch, wslot, lane, j);
	}

	float2 acc[NCH / 2][K];      // channel pairs (2c, 2c+1) x frames t + k T
	float acc17[ROOM ? K : 1];   // the reflections send
#pragma unroll
	for (int c = 0; c < NCH / 2; ++c)
#pragma unroll
		for (int k = 0; k < K; ++k) acc[c][k] = make_float2(0.0f, 0.0f);
#pragma unroll
	for (int k = 0; k < (ROOM ? K : 1); ++k) acc17[k] = 0.0f;
	const uint32_t warp_i0 = t & ~31u;     // first frame of this warp in iteration k = 0
	const float ft = (float)t;

	// acc[.][k] += gains (x) yv for one frame, constant gains: 8 FFMA2, gains straight from shared memory
	auto accumulate_const = [&](const float* gk, float yv, int k) {
		const float4* gk4 = reinterpret_cast<const float4*>(gk);
#pragma unroll
		for (int q = 0; q < NCH / 4; ++q) {
			const float4 g = gk4[q];
			ffma2(acc[2 * q][k], make_float2(g.x, g.y), make_float2(yv, yv));
			ffma2(acc[2 * q + 1][k], make_float2(g.z, g.w), make_float2(yv, yv));
		}
	};
	// same for a frame inside a ramp: channel c's gain is konst + (i < nf ? stm + i inc : 0)  (ShRamps, sbm_core.h; GainProcessor in
	// accumulate mode, gain.cc:25-92)
	auto accumulate_ramped = [&](const float* gk, const ShRamps* rp, float yv, float fi, int k) {
		const float4* gk4 = reinterpret_cast<const float4*>(gk);
#pragma unroll
		for (int q = 0; q < NCH / 4; ++q) {
			float4 g = gk4[q];
			if (rp) {     // null: no channel ramps this buffer (only the direct gain ramps), ShRamps was not staged
				const float4 nn = reinterpret_cast<const float4*>(rp->nf)[q];
				const float4 st = reinterpret_cast<const float4*>(rp->stm)[q], in = reinterpret_cast<const float4*>(rp->inc)[q];
				g.x = fi < nn.x ? g.x + fmaf(fi, in.x, st.x) : g.x; g.y = fi < nn.y ? g.y + fmaf(fi, in.y, st.y) : g.y;
				g.z = fi < nn.z ? g.z + fmaf(fi, in.z, st.z) : g.z; g.w = fi < nn.w ? g.w + fmaf(fi, in.w, st.w) : g.w;
			}
			ffma2(acc[2 * q][k], make_float2(g.x, g.y), make_float2(yv, yv));
			ffma2(acc[2 * q + 1][k], make_float2(g.z, g.w), make_float2(yv, yv));
		}
	};

	uint32_t qr_c = 0, qw_c = 0, ph_c = 0;     // (G == 1: running counters instead of two divisions per batch)
	for (uint32_t jb = grp; jb < nbatch; jb += G) {     // this group's batches
		const uint32_t qr = G == 1 ? qr_c : jb % kRecStages, qw = G == 1 ? qw_c : jb % kStages, ph_w = G == 1 ? ph_c : (jb / kStages) & 1u;     // record slot / window stage of batch jb, parity of its full_w
		if (G == 1) {
			if (++qr_c == (uint32_t)kRecStages) qr_c = 0;
			if (++qw_c == (uint32_t)kStages) { qw_c = 0; ph_c ^= 1u; }
		}
		mbar_wait(&hdr->full_w[qw], ph_w);
		const uint32_t b0 = s0 + SBM_GLOBAL_BATCH(jb) * kBatch;
		const uint32_t nb = s1 - b0 < (uint32_t)kBatch ? s1 - b0 : (uint32_t)kBatch;
		const MixRec* mrs = sm.recs + qr * kBatch;                         // this batch's records
		const float* slots = sm.wring + (size_t)(qw * kBatch) * wslot;     // ... and stage slots (source b: slots + b * wslot)
		// ---------------- whole-batch fast path ----------------
		// A batch whose sources are all "simple" (window staged by TMA, native rate or one read head, no filter, every gain ramp
		// over within the first T frames) runs in two straight-line passes:
		//   1. constant pass -- all K frames of all sources with the constant gains: per source and thread 4 sample loads, the 16
		//      gains as four LDS.128, 32 FFMA2; no branch depends on whether anything ramps;
		//   2. fix-up pass -- only the warps whose k = 0 frames lie inside a ramp, only for the sources that ramp: adds
		//      (gain(i) - konst) * y for the channel ramps, branch-free as stm + inc * min(i, nf) (ShRamps), and
		//      konst * (y(i) - y_const) for a ramping direct gain.
		bool batch_done = false;
		if (EXACT && nb == (uint32_t)kBatch) {
			uint32_t any = 0, ramps = 0, nr_max = 0;
#pragma unroll
			for (int b = 0; b < kBatch; ++b) {
				const uint2 mn = *reinterpret_cast<const uint2*>(&mrs[b]);     // {mode, nramp}
				any |= mn.x & ~(kMdInterp | kMdRamp | kMdShRamp);
				ramps |= mn.x;
				nr_max = mn.y > nr_max ? mn.y : nr_max;     // the batch's longest ramp (nramp is 0 for a source without ramps)
			}
			if (any == 0) {
				// this thread's K samples of source b, before any gain
				auto fetch = [&](int b, uint32_t w_off, float* x) {
					const float* wb = slots + (size_t)b * wslot + w_off;
					if (ramps & kMdInterp) {     // (batch-uniform test first: native-rate batches never look at the per-source mode)
						const MixRec& m = mrs[b];
						if (m.mode & kMdInterp) {
							const unsigned long long step = m.step;
							unsigned long long p = m.x0 + (unsigned long long)t * step;
							const unsigned long long stepT = step * (unsigned long long)T;
#pragma unroll
							for (int k = 0; k < K; ++k) {
								const uint32_t di = (uint32_t)(p >> 32);
								const float v0 = wb[di], v1 = wb[di + 1];
								x[k] = fmaf(frac_of(p), v1 - v0, v0);
								p += stepT;
							}
							return;
						}
					}
#pragma unroll
					for (int k = 0; k < K; ++k) x[k] = wb[t + (uint32_t)k * T];
				};
				constexpr int FUSE = TMAX == 128 ? SBM_SH_FUSE : SBM_SH_BIG_FUSE;     // sources whose sample loads are issued together
#pragma unroll
				for (int b0f = 0; b0f < kBatch; b0f += FUSE) {
					float xs[FUSE][K];
					float gds[FUSE];
#pragma unroll
					for (int bb = 0; bb < FUSE; ++bb) {
						const uint2 wg = reinterpret_cast<const uint2*>(&mrs[b0f + bb])[1];     // {w_off, gd}
						gds[bb] = __uint_as_float(wg.y);
						fetch(b0f + bb, wg.x, xs[bb]);
					}
#pragma unroll
					for (int bb = 0; bb < FUSE; ++bb) {
						const int b = b0f + bb;
						const float4* gk4 = reinterpret_cast<const float4*>(sm.shgs[qr * kBatch + b].konst);
						float2 yy[K];
#pragma unroll
						for (int k = 0; k < K; ++k) { const float y = xs[bb][k] * gds[bb]; yy[k] = make_float2(y, y); }
#pragma unroll
						for (int q = 0; q < NCH / 4; ++q) {
							const float4 g = gk4[q];
#pragma unroll
							for (int k = 0; k < K; ++k) {
								ffma2(acc[2 * q][k], make_float2(g.x, g.y), yy[k]);
								ffma2(acc[2 * q + 1][k], make_float2(g.z, g.w), yy[k]);
							}
						}
						// frame F - 1 = t + (K - 1) T of thread T - 1: the bypassed occlusion filter's memory (mono_pole_filter.cc:40-44)
						if (warp_i0 == T - 32u) { if (lane == 31u) a.lpf_prev[mrs[b].slot] = yy[K - 1].x; }     // (warp-uniform test first: only the last chunk's warp looks)
						if (ROOM) {
							// the reflections send: the raw sample (mix factor applied) through the send gain's constant part
							const float gr = sm.rfls[qr * kBatch + b].konst * mrs[b].mfx;
#pragma unroll
							for (int k = 0; k < K; ++k) acc17[k] = fmaf(gr, xs[bb][k], acc17[k]);
						}
					}
				}
				if (nr_max > warp_i0) {     // (a warp whose frames lie beyond every ramp of the batch does not even look at the sources)
#pragma unroll 1
					for (int b = 0; b < kBatch; ++b) {
						const MixRec& m = mrs[b];
						const uint32_t nramp = m.nramp;
						if (!(m.mode & kMdRamp) || warp_i0 >= nramp) continue;     // warp-uniform
						const float4* gk4 = reinterpret_cast<const float4*>(sm.shgs[qr * kBatch + b].konst);
						const ShRamps* rp = (m.mode & kMdShRamp) ? &sm.shrs[qw * kBatch + b] : nullptr;
						const bool dramp = (m.dn_nsrc & 0xffffu) != 0;
						float x[K];
						fetch(b, m.w_off, x);     // (the window is still staged: re-reading costs less than keeping the samples in registers)
#pragma unroll
						for (int k = 0; k < K; ++k) {
							// ramps usually end within the first T frames (k = 0 only); long ones (a volume change) reach further
							if (k > 0 && warp_i0 + (uint32_t)k * T >= nramp) break;
							const uint32_t i = t + (uint32_t)k * T;
							const float fi = (float)i;
							const float y0 = x[k] * m.gd;
							const float Y = (x[k] * m.mfx) * direct_gain_at(m, i);
							if (dramp) {
								const float dl = Y - y0;
#pragma unroll
								for (int q = 0; q < NCH / 4; ++q) {
									const float4 g = gk4[q];
									ffma2(acc[2 * q][k], make_float2(g.x, g.y), make_float2(dl, dl));
									ffma2(acc[2 * q + 1][k], make_float2(g.z, g.w), make_float2(dl, dl));
								}
							}
							if (rp) {
#pragma unroll
								for (int q = 0; q < NCH / 4; ++q) {
									const float4 nn = reinterpret_cast<const float4*>(rp->nf)[q];
									const float4 st = reinterpret_cast<const float4*>(rp->stm)[q], in = reinterpret_cast<const float4*>(rp->inc)[q];
									float2 e0 = make_float2(st.x, st.y), e1 = make_float2(st.z, st.w);
									ffma2(e0, make_float2(fminf(fi, nn.x), fminf(fi, nn.y)), make_float2(in.x, in.y));
									ffma2(e1, make_float2(fminf(fi, nn.z), fminf(fi, nn.w)), make_float2(in.z, in.w));
									ffma2(acc[2 * q][k], e0, make_float2(Y, Y));
									ffma2(acc[2 * q + 1][k], e1, make_float2(Y, Y));
								}
							}
							// (frame F - 1 inside a ramp: the constant pass stored y_const as the bypassed filter's memory)
							if (k == K - 1 && t == T - 1) a.lpf_prev[m.slot] = Y;
							if (ROOM) {
								const ReflSend rs = sm.rfls[qr * kBatch + b];
								if (i < rs.n) acc17[k] = fmaf(fmaf(fi, rs.inc, rs.start) - rs.konst, x[k] * m.mfx, acc17[k]);
							}
						}
					}
				}
				batch_done = true;
			}
		}
		for (uint32_t b = 0; b < nb && !batch_done; ++b) {
			const MixRec& m = mrs[b];
			const uint32_t mode = m.mode;
			if (mode & kMdSkip) {
				// gain-skipped source: the feeder still consumed this buffer; carry the resampler history forward
				if ((mode & kMdResample) && t == 0) {
					const WorkRec& r = a.rec[b0 + b];
					float p0, p1;
					resampler_tail(r, &p0, &p1);
					a.rs[r.slot].prev_samples[0] = p0; a.rs[r.slot].prev_samples[1] = p1;
				}
				continue;
			}
			float x[K];          // engine-rate input samples of this thread's frames (before any gain)
			float mf = 1.0f;     // MixSource::mix_factor where it is not already inside x
			bool filtered = false;   // x already carries the direct gain and the occlusion low-pass
			float* slot = sm.wring + (size_t)(qw * kBatch + b) * wslot;
			if ((mode & (kMdGeneric | kMdLpf | kMdAdjust)) != 0) {
				// ---------------- careful fetch (rare): every case, straightforward arithmetic, from the full record ----------------
				const WorkRec& r = a.rec[b0 + b];
				const WinPlan wp = plan_window(r, wslot);
				const float* w = slot + 4 + wp.off0;
#pragma unroll
				for (int k = 0; k < K; ++k) {
					const uint32_t i = t + (uint32_t)k * T;
					x[k] = 0.0f;
					if (i < F) x[k] = wp.fast ? staged_sample(r, w, wp, i) : source_sample(r, a.mixz, a.mix_pos0, i);
				}
				if ((r.flags & kRecResample) && t == 0) {
					float p0, p1;
					if (wp.fast) { p0 = staged_tmp(r, w, wp, (int)r.n_src); p1 = staged_tmp(r, w, wp, (int)r.n_src + 1); }
					else resampler_tail(r, &p0, &p1);
					a.rs[r.slot].prev_samples[0] = p0; a.rs[r.slot].prev_samples[1] = p1;
				}
				if (ROOM) {
					// the reflections send takes the raw samples (mix factors are inside x on this path), before gain and filter
					const ReflSend rs = sm.rfls[qr * kBatch + b];
#pragma unroll
					for (int k = 0; k < K; ++k) { const uint32_t i = t + (uint32_t)k * T; if (i < F) acc17[k] = fmaf(refl_gain_at(rs, i), x[k], acc17[k]); }
				}
				if (mode & kMdLpf) {
					// direct gain, then the recurrence in the (now free) stage slot; x becomes the filtered, gained signal
					math_sync_id(1u + grp, T);
#pragma unroll
					for (int k = 0; k < K; ++k) { const uint32_t i = t + (uint32_t)k * T; if (i < F) slot[i] = x[k] * direct_gain_at(m, i); }
					math_sync_id(1u + grp, T);
					if (t < 32) { const float carry = lpf_warp(slot, F, r.lpf_coef, r.lpf_prev, lane); if (lane == 0) a.lpf_prev[r.slot] = carry; }
					math_sync_id(1u + grp, T);
#pragma unroll
					for (int k = 0; k < K; ++k) { const uint32_t i = t + (uint32_t)k * T; if (i < F) x[k] = slot[i]; }
					fence_proxy_async();     // generic-proxy writes to the slot precede the TMA refill
					filtered = true;
				}
			} else {
				// ---------------- lean fetch: window staged in shared memory by the TMA ring ----------------
				const float* wb = slot + m.w_off;
				mf = m.mfx;
				if (mode & kMdInterp) {
					const bool rsmp = (mode & kMdResample) != 0;
#pragma unroll
					for (int k = 0; k < K; ++k) {
						const uint32_t i = t + (uint32_t)k * T;
						const uint32_t wi0 = warp_i0 + (uint32_t)k * T;
						float v = 0.0f;
						if (EXACT || i < F) {
							const unsigned long long p = m.x0 + (unsigned long long)i * m.step;
							const uint32_t di = (uint32_t)(p >> 32);
							const float fr = frac_of(p);
							float v0, v1;
							// only the first frames (resampler history) and the last ones (index clamp) need the careful form
							if (rsmp && (wi0 < 16u || wi0 + 32u > F - 1)) { v0 = staged_tmp_m(m, wb, (int)di); v1 = staged_tmp_m(m, wb, (int)di + 1); }
							else { v0 = wb[di]; v1 = wb[di + 1]; }
							v = fmaf(fr, v1 - v0, v0);
						}
						x[k] = v;
					}
					if (rsmp && t == 0) {
						const int ns = (int)(m.dn_nsrc >> 16);
						a.rs[m.slot].prev_samples[0] = staged_tmp_m(m, wb, ns);
						a.rs[m.slot].prev_samples[1] = staged_tmp_m(m, wb, ns + 1);
					}
				} else {
#pragma unroll
					for (int k = 0; k < K; ++k) {
						const uint32_t i = t + (uint32_t)k * T;
						x[k] = (EXACT || i < F) ? wb[i] : 0.0f;
					}
				}
			}
			if (ROOM && (mode & (kMdGeneric | kMdLpf | kMdAdjust)) == 0) {
				const ReflSend rs = sm.rfls[qr * kBatch + b];
#pragma unroll
				for (int k = 0; k < K; ++k) { const uint32_t i = t + (uint32_t)k * T; if (EXACT || i < F) acc17[k] = fmaf(refl_gain_at(rs, i), x[k] * mf, acc17[k]); }
			}
			// ---------------- direct gain + bus gains (GainNode, GainMixer in accumulate mode) ----------------
			const float* gk = sm.shgs[qr * kBatch + b].konst;
			const float gd = m.gd;
			const uint32_t nramp = m.nramp;     // frames in which the direct gain or a channel gain still ramps
			const ShRamps* rp = (mode & kMdShRamp) ? &sm.shrs[qw * kBatch + b] : nullptr;
#pragma unroll
			for (int k = 0; k < K; ++k) {
				const uint32_t i = t + (uint32_t)k * T;
				if (warp_i0 + (uint32_t)k * T < nramp) {
					// warp-uniform: only the warps whose frames lie inside a ramp evaluate ramps
					const float yv = filtered ? x[k] : (x[k] * mf) * direct_gain_at(m, i);
					if (!filtered && i == F - 1) a.lpf_prev[m.slot] = yv;
					accumulate_ramped(gk, rp, yv, (float)i, k);
				} else {
					const float yv = filtered ? x[k] : x[k] * gd;
					if ((k == K - 1 || !EXACT) && !filtered && i == F - 1) a.lpf_prev[m.slot] = yv;
					accumulate_const(gk, yv, k);
				}
			}
		}
		// this warp is done with window stage qw and record slot qr
		__syncwarp();
		if (SELF) {
			// the warp that finishes the batch last refills what it vacated: stage qw with batch jb + kStages, slot qr with the
			// records of batch jb + kRecStages
			uint32_t last = 0;
			if (lane == 0) {
				__threadfence_block();
				last = atomicAdd(&hdr->done[qr], 1u) == T / 32u - 1u ? 1u : 0u;
				if (last) hdr->done[qr] = 0u;
				__threadfence_block();
			}
			last = __shfl_sync(0xffffffffu, last, 0);
			if (last) {
				fill_window_stage<16, kStages, kBatch>(sm, a, s0, s1, nbatch, wslot, lane, jb + (uint32_t)kStages);
				fill_record_slot<16, kStages, kBatch>(sm, a, s0, s1, nbatch, lane, jb + (uint32_t)kRecStages);
			}
		} else if (lane == 0) { mbar_arrive(&hdr->empty_w[qw]); mbar_arrive(&hdr->empty_r[qr]); }     // hand them back to the issue warp
	}
	constexpr uint32_t ROWS = ROOM ? NCH + 1 : NCH;
	float* out = partials + ((size_t)blockIdx.x * G + grp) * ROWS * F;     // one partial bus per group
#pragma unroll
	for (int c = 0; c < NCH / 2; ++c)
#pragma unroll
		for (int k = 0; k < K; ++k) {
			const uint32_t i = t + (uint32_t)k * T;
			if (EXACT || i < F) { out[(size_t)(2 * c) * F + i] = acc[c][k].x; out[(size_t)(2 * c + 1) * F + i] = acc[c][k].y; }
		}
	if (ROOM) {
#pragma unroll
		for (int k = 0; k < K; ++k) { const uint32_t i = t + (uint32_t)k * T; if (EXACT || i < F) out[(size_t)NCH * F + i] = acc17[k]; }
	}
}

// Entry points of the ambisonic kernel.  frames 256 / 512 / 1024 (T = frames / 4 = 64 / 128 / 256 math threads, no per-frame bounds
// checks): measured on B200 per shape --
//   256:  all-math 2-warp CTAs with self-service rings (8 CTAs per SM): 0.030 ms per 33 792 sources vs 0.040 with an issue warp;
//   512:  4 math warps + issue warp, 3 CTAs per SM: 0.22 ms per 262 144 sources.  The all-math form (4 CTAs x 4 warps = 16 math
//         warps) measured 0.29: the warp that owns the first 32 frames does the ramp fix-ups of every source, finishes every
//         batch last, and then also has to refill the rings while the other warps wait for it;
//   1024: 8 math warps + issue warp, 1 CTA per SM (a 9-warp CTA at > 102 registers cannot be resident twice: the register file
//         is per scheduler, 4 x 16 K): 0.150 ms per 65 536 sources; all-math 2 x 8 warps: 0.157.
// Any other frame count: the issue-warp form with per-frame bounds checks.
#ifndef SBM_STAGES_SELF
#define SBM_STAGES_SELF 4
#endif
#ifndef SBM_SH_SELF_MASK
#define SBM_SH_SELF_MASK 64           // OR of the T values that use the self-service form (A/B: 64 | 128 | 256)
#endif
#ifndef SBM_STAGES_SHX
#define SBM_STAGES_SHX 6
#endif
#ifndef SBM_STAGES_SHX_BIG
#define SBM_STAGES_SHX_BIG 3          // (x 8 sources per stage: the bytes in flight of 6 stages of 4)
#endif
constexpr int kStagesSelf = SBM_STAGES_SELF, kStagesShX = SBM_STAGES_SHX;
__host__ __device__ constexpr bool sh_self_shape(int T) { return (SBM_SH_SELF_MASK & T) != 0; }
// sources per stage / window stages of the exact shape with T math threads
__host__ __device__ constexpr int shx_batch(int T) { return T == 256 && !sh_self_shape(T) ? SBM_BATCH_SH_BIG : kBatchSh; }
__host__ __device__ constexpr int shx_stages(int T) { return sh_self_shape(T) ? kStagesSelf : (T == 256 ? SBM_STAGES_SHX_BIG : kStagesShX); }
template <int TMAX, bool ROOM>
__global__ void __launch_bounds__(TMAX + (sh_self_shape(TMAX) ? 0 : 32), sh_self_shape(TMAX) ? 512 / TMAX : (TMAX == 256 ? 1 : (TMAX == 128 ? 3 : 5)))
k2_mix_ambisonic_x_kernel(MixArgs a, uint32_t first, uint32_t n, uint32_t wslot, float* __restrict__ partials) {
	ambisonic_body<TMAX, kShK, 1, true, sh_self_shape(TMAX), shx_stages(TMAX), ROOM, shx_batch(TMAX)>(a, first, n, wslot, partials);
}
template <bool ROOM>
__global__ void __launch_bounds__(128 + 32, 3) k2_mix_ambisonic_kernel(MixArgs a, uint32_t first, uint32_t n, uint32_t wslot, float* __restrict__ partials) {
	ambisonic_body<128, kShK, 1, false, false, kStagesSh, ROOM>(a, first, n, wslot, partials);
}
template <bool ROOM>
__global__ void __launch_bounds__(kShBigT + 32, 1) k2_mix_ambisonic_big_kernel(MixArgs a, uint32_t first, uint32_t n, uint32_t wslot, float* __restrict__ partials) {
	ambisonic_body<kShBigT, kShBigK, 1, false, false, kStagesShBig, ROOM>(a, first, n, wslot, partials);
}

// K3: bus[row][f] = sum over CTAs of the partial buses; rows 0-15 from the ambisonic kernel's partials, rows 16-17 from
// the stereo kernel's.  One CTA per (row, 32 frames): warp w sums the partials of CTAs w, w+32, w+64, ... and the 32
// warp sums are added in warp order -- a fixed tree, so the result is bit-stable from run to run (no atomics on audio).
// Also publishes this rank's "mixer used" flags as floats behind the audio so that they sum across ranks with it.
constexpr uint32_t kK3Warps = 32;
__global__ void __launch_bounds__(kK3Warps * 32) k3_reduce_kernel(const float* __restrict__ part_pan, uint32_t g_pan, const float* __restrict__ part_sh,
                                                       uint32_t g_sh, float* __restrict__ bus, const DevFlags* __restrict__ flags,
                                                       uint32_t parity, uint32_t F, PeerArgs pa, size_t bus_floats, uint32_t room) {
	__shared__ float warp_sum[kK3Warps][32];
	grid_dep_wait();
	const uint32_t row = blockIdx.y, lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
	const uint32_t f = blockIdx.x * 32 + lane;
	// partial buses: [16 (+1)][F] per ambisonic CTA, [2 (+1)][F] per stereo-bus CTA; the extra row is the reflections send
	// (room effects), whose bus row (kRowRefl) sums BOTH kernels' extra rows
	const uint32_t rows_sh = room ? 17u : 16u, rows_pan = room ? 3u : 2u;
	float s = 0.0f;
	auto sum_partials = [&](const float* part, uint32_t g, uint32_t rows, uint32_t r) {
		// four independent running sums (fixed assignment b -> sum, fixed combination order: still bit-stable) keep
		// several loads in flight per lane instead of one dependent add chain
		float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
		const float* pp = part + (size_t)r * F + f;
		const size_t stride = (size_t)rows * F;
		uint32_t b = w;
		for (; b + 3 * kK3Warps < g; b += 4 * kK3Warps) {
			s0 += pp[(size_t)b * stride]; s1 += pp[(size_t)(b + kK3Warps) * stride]; s2 += pp[(size_t)(b + 2 * kK3Warps) * stride]; s3 += pp[(size_t)(b + 3 * kK3Warps) * stride];
		}
		for (; b < g; b += kK3Warps) s0 += pp[(size_t)b * stride];
		return (s0 + s1) + (s2 + s3);
	};
	if (f < F) {
		if (row < 16) s = sum_partials(part_sh, g_sh, rows_sh, row);
		else if (row < 18) s = sum_partials(part_pan, g_pan, rows_pan, row - 16);
		else if (room) s = sum_partials(part_sh, g_sh, rows_sh, 16) + sum_partials(part_pan, g_pan, rows_pan, 2);
	}
	warp_sum[w][lane] = s;
	__syncthreads();
	if (w == 0 && f < F) {
		float t = warp_sum[0][lane];
#pragma unroll
		for (int q = 1; q < (int)kK3Warps; ++q) t += warp_sum[q][lane];
		bus[(size_t)row * F + f] = t;
	}
	if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
		bus[(size_t)kBusRows * F + 0] = flags->sh_used[parity] ? 1.0f : 0.0f;
		bus[(size_t)kBusRows * F + 1] = flags->pan_used[parity] ? 1.0f : 0.0f;
		bus[(size_t)kBusRows * F + 2] = flags->st_used[parity] ? 1.0f : 0.0f;
		bus[(size_t)kBusRows * F + 3] = flags->refl_used[parity] ? 1.0f : 0.0f;
	}
	if (pa.world) {
		// Multi-GPU exchange.  Every CTA publishes its piece of the bus device-wide and counts itself; the CTA that completes the count
		// -- ONE CTA, all of its threads -- copies the whole bus (39 KB) into this rank's slot of every rank's gather area over NVLink,
		// makes those stores visible system-wide with ONE fence, and signals one arrival per rank.  (Until the end of round 2 every CTA
		// stored its own piece to the peers: 300 CTAs x a system-scope fence each, 0.031 ms per buffer on top of waiting for the
		// slowest rank; now 0.014 with the acquire loads of k_peer_gather.  Folding the wait-and-sum of k_peer_gather into this CTA as well measured 0.003 ms slower: one CTA
		// summing behind a spin is no match for 39 blocks that start under programmatic dependent launch.)
		__shared__ int is_last;
		__threadfence();
		__syncthreads();
		if (threadIdx.x == 0) {
			const unsigned int total = gridDim.x * gridDim.y;
			const bool last = atomicAdd(pa.done, 1u) == total - 1u;
			if (last) { *pa.done = 0u; __threadfence(); }
			is_last = last ? 1 : 0;
		}
		__syncthreads();
		if (is_last) {
			const size_t slot_off = ((size_t)pa.parity * pa.world + pa.rank) * bus_floats;     // this rank's slot in rank p's gather area
			for (size_t i = threadIdx.x; i < bus_floats; i += blockDim.x) {
				const float v = __ldcg(bus + i);
				for (uint32_t p = 0; p < pa.world; ++p) pa.gather[p][slot_off + i] = v;     // NVLink peer stores (local for p == rank)
			}
			__threadfence_system();
			__syncthreads();
			if (threadIdx.x < pa.world) atomicAdd_system(pa.arrive[threadIdx.x] + pa.parity, 1u);
		}
	}
}

// Gather kernel in front of K4 (multi-GPU): waits until all `world` ranks have delivered this buffer's bus into the
// local gather area, then sums the slots in rank order -- every rank computes the identical sum.  A peer that does not
// arrive within pa.timeout_ns (stalled or crashed rank) does not take this rank down: the kernel flags the buffer
// (DevFlags::peer_timeout -> OutTail::status -> SBM_ERR_PEER_TIMEOUT) and renders this rank's own bus alone.
__device__ __forceinline__ unsigned long long global_timer_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}
__global__ void __launch_bounds__(256) k_peer_gather_kernel(PeerArgs pa, const float* __restrict__ own_bus, float* __restrict__ bus_sum, size_t bus_floats,
                                                            unsigned int target, DevFlags* flags) {
	__shared__ int timed_out;
	grid_dep_wait();
	if (threadIdx.x == 0) {
		// (acquire at system scope on the counter the peers bump after their system-scope fence: no fence of our own)
		const unsigned int* cnt = pa.arrive[pa.rank] + pa.parity;
		const unsigned long long t0 = global_timer_ns();
		int late = 0;
		for (;;) {
			unsigned int v;
			asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
			if ((int)(v - target) >= 0) break;
			__nanosleep(32);
			if (global_timer_ns() - t0 > pa.timeout_ns) { late = 1; break; }
		}
		timed_out = late;
		if (late && blockIdx.x == 0) flags->peer_timeout = 1;
	}
	__syncthreads();
	const float* area = pa.gather[pa.rank] + (size_t)pa.parity * pa.world * bus_floats;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < bus_floats; i += (size_t)gridDim.x * blockDim.x) {
		float s = 0.0f;
		if (timed_out) s = own_bus[i];
		else for (uint32_t p = 0; p < pa.world; ++p) s += __ldcg(area + (size_t)p * bus_floats + i);
		bus_sum[i] = s;
	}
}

// Everything a launch needs to know about the CURRENT device, cached per device ordinal (engines on different devices and
// threads share this library): SM count, and per kernel the dynamic shared memory it was configured for and its occupancy.
struct KernelConfig { size_t smem = 0; int ctas_per_sm = 0; };
struct DeviceCache { int sms = 0; std::map<const void*, KernelConfig> kernels; };
std::mutex g_cfg_mu;
std::map<int, DeviceCache> g_devices;

DeviceCache& device_cache_locked() {
	int dev = 0;
	cudaGetDevice(&dev);
	DeviceCache& c = g_devices[dev];
	if (!c.sms) {
		cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev);
		if (c.sms <= 0) c.sms = 148;
	}
	return c;
}

int sm_count() {
	std::lock_guard<std::mutex> lk(g_cfg_mu);
	return device_cache_locked().sms;
}

uint32_t sh_frames_per_thread(uint32_t F) { return F <= 512 ? (uint32_t)kShK : (uint32_t)kShBigK; }
uint32_t sh_threads(uint32_t F) { const uint32_t K = sh_frames_per_thread(F); return ((F + K - 1) / K + 31) / 32 * 32; }
bool sh_exact(uint32_t F) { return F == 256 || F == 512 || F == 1024; }     // the exact shapes (T = F / 4 = 64 / 128 / 256 math threads)
int mix_stages(uint32_t F, bool sh) { return !sh ? kStagesPan : (sh_exact(F) ? shx_stages((int)F / 4) : (sh_threads(F) <= 128 ? kStagesSh : kStagesShBig)); }
size_t mix_smem_bytes(uint32_t F, bool sh, bool room = false, int stages = 0) {
	const size_t st = (size_t)(stages ? stages : mix_stages(F, sh)), rst = st + 2, kBatch = sh ? (size_t)(sh_exact(F) ? shx_batch((int)F / 4) : kBatchSh) : kBatchPan;
	return 512 + sizeof(MixRec) * rst * kBatch + (room ? sizeof(ReflSend) * rst * kBatch : 0) + (sh ? sizeof(ShGains) * rst * kBatch + sizeof(ShRamps) * st * kBatch : 0) +
	       sizeof(float) * st * kBatch * window_slot(F);
}
constexpr uint32_t kMinSourcesPerCta = 16;

uint32_t pick_grid(uint32_t n, uint32_t max_grid) {
	uint32_t g = (n + kMinSourcesPerCta - 1) / kMinSourcesPerCta;
	if (g > max_grid) g = max_grid;
	return g ? g : 1;
}

// Launches one K2 kernel as a single wave: grid = SM count x resident CTAs per SM (never more than the sources warrant).
template <class Kernel>
int launch_one(Kernel kernel, const char* name, const MixArgs& a, uint32_t first, uint32_t n, uint32_t threads, bool sh,
               uint32_t max_grid, float* partials, cudaStream_t st, uint32_t* grid_out, int stages = 0) {
	const uint32_t F = a.frames;
	const size_t smem = mix_smem_bytes(F, sh, a.rfl != nullptr, stages);
	uint32_t wave;
	int sm_count_now = 148;
	{
		// cudaFuncSetAttribute and the occupancy query apply to the current device only: configured once per (device, kernel, size)
		std::lock_guard<std::mutex> lk(g_cfg_mu);
		DeviceCache& dc = device_cache_locked();
		KernelConfig* cfg = &dc.kernels[(const void*)kernel];
		if (smem != cfg->smem) {
			if (dev::check(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(smem)") ||
			    dev::check(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), "cudaFuncSetAttribute(carveout)"))
				return -1;
			int per_sm = 0;
			if (dev::check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)threads, smem), "cudaOccupancyMaxActiveBlocksPerMultiprocessor")) return -1;
			cfg->smem = smem;
			cfg->ctas_per_sm = per_sm > 0 ? per_sm : 1;
		}
		wave = (uint32_t)(dc.sms * cfg->ctas_per_sm);
		sm_count_now = dc.sms;
	}
	if (wave > max_grid) wave = max_grid;
	const uint32_t g = pick_grid(n, wave);
	MixArgs a2 = a;
	a2.rot_div = (uint32_t)sm_count_now;
	launch_pdl(kernel, dim3(g), dim3(threads), smem, st, a2, first, n, window_slot(F), partials);
	dev::note_launch();
	*grid_out = g;
	return dev::check_launch(name);
}

template <int KF>
int launch_stereo(const MixArgs& a, uint32_t max_grid, float* partials, cudaStream_t st, uint32_t* grid_out) {
	const uint32_t threads = kPanWarps * 32 + 32;
	if (a.rfl) {
		if (a.frames == 32u * KF) return launch_one(k2_mix_stereo_kernel<KF, true, true, kStagesPan>, "k2_mix_stereo", a, 0, a.n_pan, threads, false, max_grid, partials, st, grid_out);
		return launch_one(k2_mix_stereo_kernel<KF, false, true, kStagesPan>, "k2_mix_stereo", a, 0, a.n_pan, threads, false, max_grid, partials, st, grid_out);
	}
	if (a.ready) {      // K1 is running beside this launch: the shape that leaves it room
		if (a.frames == 32u * KF) return launch_one(k2_mix_stereo_kernel<KF, true, false, kStagesPanOverlap>, "k2_mix_stereo", a, 0, a.n_pan, threads, false, max_grid, partials, st, grid_out, kStagesPanOverlap);
		return launch_one(k2_mix_stereo_kernel<KF, false, false, kStagesPanOverlap>, "k2_mix_stereo", a, 0, a.n_pan, threads, false, max_grid, partials, st, grid_out, kStagesPanOverlap);
	}
	if (a.frames == 32u * KF) return launch_one(k2_mix_stereo_kernel<KF, true, false, kStagesPan>, "k2_mix_stereo", a, 0, a.n_pan, threads, false, max_grid, partials, st, grid_out);
	return launch_one(k2_mix_stereo_kernel<KF, false, false, kStagesPan>, "k2_mix_stereo", a, 0, a.n_pan, threads, false, max_grid, partials, st, grid_out);
}

}  // namespace

namespace dev {

int current_sm_count() { return sm_count(); }

// Upper bound of the K2 grid (sizes the partial-bus scratch): SM count x the most CTAs that fit per SM.
int num_mix_ctas(uint32_t frames, int sh) {
	const size_t smem = mix_smem_bytes(frames, sh != 0) + 1024;
	int per_sm = (int)((227u * 1024u) / smem);
	if (per_sm < 1) per_sm = 1;
	if (per_sm > 8) per_sm = 8;
	return sm_count() * per_sm;
}

size_t partials_floats(uint32_t frames) {
	return ((size_t)num_mix_ctas(frames, 0) * 3 + (size_t)num_mix_ctas(frames, 1) * 17) * frames;     // (+1 row each: the reflections send, room effects)
}

int launch_mix(const MixArgs& a, void* stream) {
	cudaStream_t st = (cudaStream_t)stream;
	const uint32_t F = a.frames;
	float* part_pan = a.partials;
	float* part_sh = a.partials + (size_t)num_mix_ctas(F, 0) * 3 * F;
	uint32_t g_pan = 0, g_sh = 0;
	if (a.n_pan) {
		const uint32_t mg = (uint32_t)num_mix_ctas(F, 0);
		int rc;
		if (F <= 256) rc = launch_stereo<8>(a, mg, part_pan, st, &g_pan);
		else if (F <= 512) rc = launch_stereo<16>(a, mg, part_pan, st, &g_pan);
		else rc = launch_stereo<32>(a, mg, part_pan, st, &g_pan);
		if (rc) return -1;
	}
	if (a.n_sh) {
		const uint32_t tm = sh_threads(F);
		const uint32_t mg = (uint32_t)num_mix_ctas(F, 1);
		int rc;
		const bool room = a.rfl != nullptr;
#define SBM_LAUNCH_SH(KERNEL, THREADS) (room ? launch_one(KERNEL<true>, "k2_mix_ambisonic", a, a.n_pan, a.n_sh, THREADS, true, mg, part_sh, st, &g_sh) \
                                             : launch_one(KERNEL<false>, "k2_mix_ambisonic", a, a.n_pan, a.n_sh, THREADS, true, mg, part_sh, st, &g_sh))
#define SBM_LAUNCH_SHX(TT) (room ? launch_one(k2_mix_ambisonic_x_kernel<TT, true>, "k2_mix_ambisonic", a, a.n_pan, a.n_sh, TT + (sh_self_shape(TT) ? 0 : 32), true, mg, part_sh, st, &g_sh) \
                                 : launch_one(k2_mix_ambisonic_x_kernel<TT, false>, "k2_mix_ambisonic", a, a.n_pan, a.n_sh, TT + (sh_self_shape(TT) ? 0 : 32), true, mg, part_sh, st, &g_sh))
		if (sh_exact(F)) {
			if (F == 256) rc = SBM_LAUNCH_SHX(64);
			else if (F == 512) rc = SBM_LAUNCH_SHX(128);
			else rc = SBM_LAUNCH_SHX(256);
		} else if (tm <= 128) rc = SBM_LAUNCH_SH(k2_mix_ambisonic_kernel, tm + 32);
		else rc = SBM_LAUNCH_SH(k2_mix_ambisonic_big_kernel, tm + 32);
#undef SBM_LAUNCH_SH
#undef SBM_LAUNCH_SHX
		if (rc) return -1;
	}
	if (a.ev_after_k2) cudaEventRecord((cudaEvent_t)a.ev_after_k2, st);
	launch_pdl(k3_reduce_kernel, dim3((F + 31) / 32, kBusRows), dim3(kK3Warps * 32), 0, st, (const float*)part_pan, g_pan, (const float*)part_sh, g_sh, a.bus, a.flags, a.parity, F, a.peer, a.bus_floats, a.rfl ? 1u : 0u);
	note_launch();
	return check_launch("k3_reduce");
}

int launch_peer_gather(const PeerArgs& pa, const float* own_bus, float* bus_sum, size_t bus_floats, unsigned int target, DevFlags* flags, void* stream) {
	const uint32_t grid = (uint32_t)((bus_floats + 255) / 256);
	launch_pdl(k_peer_gather_kernel, dim3(grid < 148u ? grid : 148u), dim3(256), 0, (cudaStream_t)stream, pa, own_bus, bus_sum, bus_floats, target, flags);
	note_launch();
	return check_launch("k_peer_gather");
}

}  // namespace dev
}  // namespace sbm
