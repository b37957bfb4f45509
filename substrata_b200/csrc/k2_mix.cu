// k2_mix.cu -- "K2": the hot kernel.  For every rendered source of this buffer: gather its sample window from HBM
// (loop / ring wrap, linear sample-rate conversion, pitch / doppler read heads), apply the ramped direct gain and the
// optional occlusion low-pass, multiply by the source's 2 (stereo pan / non-spatial) or 16 (3rd-order ambisonic)
// ramped bus gains and accumulate into the output tile; and "K3", the fixed-order cross-CTA reduction of the partial
// buses into this rank's bus.
//
// Decomposition: a CTA owns ALL frames of the buffer and a contiguous slice of the source list.  Thread t owns frames
// t, t+T, t+2T, ... (T = blockDim.x), so that for one source a warp reads 32 consecutive samples per load (coalesced
// 128 B) and every thread keeps its NCH x K partial sums in registers across the CTA's whole slice of sources -- the
// per-source contributions never leave the register file until the slice is done.  Each CTA then stores one partial
// bus [NCH][frames]; K3 sums the partial buses in CTA order (deterministic; no atomics on audio).
// HBM traffic: 4 B per source sample read (PCM) + 128 B WorkRec (+64 B SH gains) per source; writes are O(grid * frames).
// Roofline: HBM read.  No tensor cores: 2..16 FMA per 4 B loaded, no contraction to feed them.
#include "device_internal.h"
#include "k2_core.h"

namespace sbm {

namespace {

constexpr int kBatch = 16;            // WorkRecs staged in shared memory per step
constexpr int kPanThreads = 128;
constexpr int kShK = 4;               // frames per thread in the 16-channel kernel

struct MixSmem {
	WorkRec rec[kBatch];
	ShGains shg[kBatch];
	float lpf[kMaxFrames];
	float lpf_carry;
};

// Occlusion low-pass over the CTA's staged buffer (rare path: only occluded sources).  y[n] = c (y[n-1] - x[n]) + x[n]
// is linear in y[n-1]: warp 0 splits the buffer into 32 contiguous chunks, runs each chunk from a zero state while
// accumulating c^len, chains the 32 chunk states in order, and re-runs each chunk from its true initial state.
__device__ void lpf_in_smem(float* buf, uint32_t F, float coef, float prev, float* carry_out) {
	if (threadIdx.x < 32) {
		const uint32_t lane = threadIdx.x;
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
		if (lane == 0) *carry_out = carry;
	}
}

template <int NCH, int K>
__global__ void __launch_bounds__(NCH == 2 ? kPanThreads : 256)
k2_mix_kernel(MixArgs a, const uint32_t* __restrict__ list, uint32_t n, float* __restrict__ partials) {
	__shared__ MixSmem sm;
	const uint32_t F = a.frames, T = blockDim.x, t = threadIdx.x;
	const uint32_t s0 = (uint32_t)((unsigned long long)blockIdx.x * n / gridDim.x);
	const uint32_t s1 = (uint32_t)((unsigned long long)(blockIdx.x + 1) * n / gridDim.x);
	float acc[NCH][K];
#pragma unroll
	for (int c = 0; c < NCH; ++c)
#pragma unroll
		for (int k = 0; k < K; ++k) acc[c][k] = 0.0f;

	for (uint32_t base = s0; base < s1; base += kBatch) {
		const uint32_t nb = s1 - base < (uint32_t)kBatch ? s1 - base : (uint32_t)kBatch;
		__syncthreads();
		// stage the batch's WorkRecs (8 x 16 B each) and, for ambisonic sources, their 16 constant gains (4 x 16 B)
		for (uint32_t q = t; q < nb * 8; q += T) {
			const uint32_t slot = list[base + (q >> 3)];
			reinterpret_cast<uint4*>(sm.rec)[q] = __ldg(reinterpret_cast<const uint4*>(a.rec + slot) + (q & 7));
		}
		if (NCH == 16)
			for (uint32_t q = t; q < nb * 4; q += T) {
				const uint32_t slot = list[base + (q >> 2)];
				reinterpret_cast<uint4*>(sm.shg)[q] = __ldg(reinterpret_cast<const uint4*>(a.shg + slot) + (q & 3));
			}
		__syncthreads();

		for (uint32_t j = 0; j < nb; ++j) {
			const WorkRec& r = sm.rec[j];
			const uint32_t flags = r.flags;
			if (flags & kRecSkip) {
				// gain-skipped source: the feeder still consumed this buffer; carry the resampler history forward
				if ((flags & kRecResample) && t == 0) {
					float p0, p1;
					resampler_tail(r, &p0, &p1);
					a.rs[r.slot].prev_samples[0] = p0; a.rs[r.slot].prev_samples[1] = p1;
				}
				continue;
			}
			const double w = (double)r.rate / (double)a.rate;
			float y[K];
			// ---- gather / resample / pitch-mix, then the direct gain (GainNode) ----
#pragma unroll
			for (int k = 0; k < K; ++k) {
				const uint32_t i = t + (uint32_t)k * T;
				y[k] = i < F ? source_sample(r, a.mixz, a.mix_pos0, w, i) : 0.0f;
			}
			if (r.direct.n) {
#pragma unroll
				for (int k = 0; k < K; ++k) y[k] *= gain_at(r.direct, t + (uint32_t)k * T);
			} else {
				const float g = r.direct.konst;
#pragma unroll
				for (int k = 0; k < K; ++k) y[k] *= g;
			}
			if ((flags & kRecResample) && t == 0) {
				float p0, p1;
				resampler_tail(r, &p0, &p1);
				a.rs[r.slot].prev_samples[0] = p0; a.rs[r.slot].prev_samples[1] = p1;
			}
			// ---- occlusion low-pass (OcclusionNode) ----
			if (!(flags & kRecStereo)) {
				if (flags & kRecLpf) {
					__syncthreads();
#pragma unroll
					for (int k = 0; k < K; ++k) { const uint32_t i = t + (uint32_t)k * T; if (i < F) sm.lpf[i] = y[k]; }
					__syncthreads();
					lpf_in_smem(sm.lpf, F, r.lpf_coef, r.lpf_prev, &sm.lpf_carry);
					__syncthreads();
#pragma unroll
					for (int k = 0; k < K; ++k) { const uint32_t i = t + (uint32_t)k * T; if (i < F) y[k] = sm.lpf[i]; }
					if (t == 0) a.lpf_prev[r.slot] = sm.lpf_carry;
				} else {
					// bypassed filter remembers the last input sample (mono_pole_filter.cc:40-44)
#pragma unroll
					for (int k = 0; k < K; ++k) if (t + (uint32_t)k * T == F - 1) a.lpf_prev[r.slot] = y[k];
				}
			}
			// ---- bus gains (GainMixer, accumulate mode) ----
			if (NCH == 2) {
				if (r.ch_ramp_mask) {
#pragma unroll
					for (int k = 0; k < K; ++k) {
						const uint32_t i = t + (uint32_t)k * T;
						acc[0][k] = fmaf(gain_at(r.ch[0], i), y[k], acc[0][k]);
						acc[1][k] = fmaf(gain_at(r.ch[1], i), y[k], acc[1][k]);
					}
				} else {
					const float g0 = r.ch[0].konst, g1 = r.ch[1].konst;
#pragma unroll
					for (int k = 0; k < K; ++k) { acc[0][k] = fmaf(g0, y[k], acc[0][k]); acc[1][k] = fmaf(g1, y[k], acc[1][k]); }
				}
			} else {
				const float* gk = sm.shg[j].konst;
				if (r.ch_ramp_mask) {
					const ShRamps* rp = a.shr + r.slot;
#pragma unroll
					for (int c = 0; c < NCH; ++c) {
						RampDesc d; d.konst = gk[c]; d.n = 0; d.start = 0.0f; d.inc = 0.0f;
						if (r.ch_ramp_mask & (1u << c)) { d.start = rp->start[c]; d.inc = rp->inc[c]; d.n = rp->n[c]; }
#pragma unroll
						for (int k = 0; k < K; ++k) acc[c][k] = fmaf(gain_at(d, t + (uint32_t)k * T), y[k], acc[c][k]);
					}
				} else {
#pragma unroll
					for (int c = 0; c < NCH; ++c) {
						const float g = gk[c];
#pragma unroll
						for (int k = 0; k < K; ++k) acc[c][k] = fmaf(g, y[k], acc[c][k]);
					}
				}
			}
		}
	}
	float* out = partials + (size_t)blockIdx.x * NCH * F;
#pragma unroll
	for (int c = 0; c < NCH; ++c)
#pragma unroll
		for (int k = 0; k < K; ++k) {
			const uint32_t i = t + (uint32_t)k * T;
			if (i < F) out[(size_t)c * F + i] = acc[c][k];
		}
}

// K3: bus[row][f] = sum over CTAs (ascending) of the partial buses; rows 0-15 from the ambisonic kernel's partials,
// rows 16-17 from the stereo kernel's.  Also publishes this rank's "mixer used" flags as floats behind the audio so
// that they sum across ranks with it.
__global__ void __launch_bounds__(256) k3_reduce_kernel(const float* __restrict__ part_pan, uint32_t g_pan, const float* __restrict__ part_sh,
                                                       uint32_t g_sh, float* __restrict__ bus, const DevFlags* __restrict__ flags,
                                                       uint32_t parity, uint32_t F) {
	const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx < kBusRows * F) {
		const uint32_t row = idx / F, f = idx - row * F;
		float s = 0.0f;
		if (row < 16) for (uint32_t b = 0; b < g_sh; ++b) s += part_sh[((size_t)b * 16 + row) * F + f];
		else for (uint32_t b = 0; b < g_pan; ++b) s += part_pan[((size_t)b * 2 + (row - 16)) * F + f];
		bus[idx] = s;
	}
	if (idx == 0) {
		float* bf = bus + (size_t)kBusRows * F;
		bf[0] = flags->sh_used[parity] ? 1.0f : 0.0f;
		bf[1] = flags->pan_used[parity] ? 1.0f : 0.0f;
		bf[2] = flags->st_used[parity] ? 1.0f : 0.0f;
		bf[3] = 0.0f;
	}
}

int sm_count() {
	static int n = 0;
	if (!n) {
		int dev = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
		if (n <= 0) n = 148;
	}
	return n;
}

uint32_t sh_threads(uint32_t F) { return ((F + kShK - 1) / kShK + 31) / 32 * 32; }
constexpr uint32_t kMinSourcesPerCta = 8;

uint32_t pick_grid(uint32_t n, uint32_t max_grid) {
	uint32_t g = (n + kMinSourcesPerCta - 1) / kMinSourcesPerCta;
	if (g > max_grid) g = max_grid;
	return g ? g : 1;
}

}  // namespace

namespace dev {

int num_mix_ctas(uint32_t frames, int sh) {
	(void)frames;
	return sm_count() * (sh ? 4 : 8);
}

size_t partials_floats(uint32_t frames) {
	return ((size_t)num_mix_ctas(frames, 0) * 2 + (size_t)num_mix_ctas(frames, 1) * 16) * frames;
}

int launch_mix(const MixArgs& a, void* stream) {
	cudaStream_t st = (cudaStream_t)stream;
	const uint32_t F = a.frames;
	float* part_pan = a.partials;
	float* part_sh = a.partials + (size_t)num_mix_ctas(F, 0) * 2 * F;
	uint32_t g_pan = 0, g_sh = 0;
	if (a.n_pan) {
		g_pan = pick_grid(a.n_pan, (uint32_t)num_mix_ctas(F, 0));
		const uint32_t k = (F + kPanThreads - 1) / kPanThreads;
		if (k <= 1) k2_mix_kernel<2, 1><<<g_pan, kPanThreads, 0, st>>>(a, a.list_pan, a.n_pan, part_pan);
		else if (k <= 2) k2_mix_kernel<2, 2><<<g_pan, kPanThreads, 0, st>>>(a, a.list_pan, a.n_pan, part_pan);
		else if (k <= 4) k2_mix_kernel<2, 4><<<g_pan, kPanThreads, 0, st>>>(a, a.list_pan, a.n_pan, part_pan);
		else k2_mix_kernel<2, 8><<<g_pan, kPanThreads, 0, st>>>(a, a.list_pan, a.n_pan, part_pan);
		note_launch();
		if (check_launch("k2_mix<pan>")) return -1;
	}
	if (a.n_sh) {
		g_sh = pick_grid(a.n_sh, (uint32_t)num_mix_ctas(F, 1));
		k2_mix_kernel<16, kShK><<<g_sh, sh_threads(F), 0, st>>>(a, a.list_sh, a.n_sh, part_sh);
		note_launch();
		if (check_launch("k2_mix<sh>")) return -1;
	}
	if (a.ev_after_k2) cudaEventRecord((cudaEvent_t)a.ev_after_k2, st);
	k3_reduce_kernel<<<(kBusRows * F + 255) / 256, 256, 0, st>>>(part_pan, g_pan, part_sh, g_sh, a.bus, a.flags, a.parity, F);
	note_launch();
	return check_launch("k3_reduce");
}

}  // namespace dev
}  // namespace sbm
