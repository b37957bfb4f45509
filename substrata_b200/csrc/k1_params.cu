// k1_params.cu -- "K1": one thread per render-list entry (live source) runs the feeder bookkeeping and Resonance's per-source parameter
// logic for this buffer (k1_core.h: update_source) and condenses the slot into one WorkRec for the mix kernel; plus
// k_patch, which applies host-requested state (re)initialisations before K1.
//
// Compiled with -fmad=false: results of this unit feed integer truncations and thresholds (lookup-table cell,
// -60 dB rules, ramp lengths), so every float operation must round exactly like the reference's SSE build.
// HBM traffic per slot: SrcStatic 32 B + pos/volume 16 B + ~60 B of state read-modify-write + 128 B MixRec write (k2_rec.h;
// + the 128 B WorkRec only for sources a careful path of the mix kernel will read back).
#include "device_internal.h"
#include "k1_core.h"
#include "k2_rec.h"

namespace sbm {

#ifndef SBM_K1_THREADS
#define SBM_K1_THREADS 128             // 64 .. 256 measure within 3 %; 512 is 5 % slower (one thread per source, latency-bound)
#endif
#ifndef SBM_K1_MINB
#define SBM_K1_MINB (1024 / SBM_K1_THREADS)
#endif
__global__ void __launch_bounds__(SBM_K1_THREADS, SBM_K1_MINB) k1_params_kernel(Globals g, SlotArrays A, DevFlags* flags, const float* __restrict__ table,
                                                       const uint32_t* __restrict__ lists, uint32_t wslot) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;    // position in the render lists
	// whether this buffer's "mixer used" flags are already set: read NOW, so that the loads overlap the work below instead
	// of being a dependent round trip at the end of every warp (a stale 0 only costs a redundant atomicOr)
	const int set_pan = flags->pan_used[g.parity], set_sh = flags->sh_used[g.parity], set_st = flags->st_used[g.parity], set_refl = flags->refl_used[g.parity];
	uint32_t used = 0, n_src = 0;
	if (k < g.n_pan + g.n_sh) {
		WorkRec rec;
		ReflSend rfl;
		rfl.konst = 0.0f; rfl.start = 0.0f; rfl.inc = 0.0f; rfl.n = 0;
		used = update_source(g, A, flags, table, lists[k], &rec, A.shg + k, A.shr + k, &rfl, &n_src);
		// the record the mix kernel consumes, in final form (k2_rec.h); the full WorkRec only where a careful path will read it back
		MixRec m = make_mixrec(rec, wslot, k >= g.n_pan);
		if (g.room) {
			// a ramping reflections send needs the per-frame gain path like any other ramp
			A.rfl[k] = rfl;
			if (rfl.n && !(m.mode & kMdSkip)) { if (rfl.n > m.nramp) m.nramp = rfl.n; m.mode |= kMdRamp; }
		}
		A.mrec[k] = m;
		if (mixrec_needs_workrec(m.mode)) A.rec[k] = rec;
	}
	// Warp-level reduction of the "mixer used" bits and the accounting counters, then one lane publishes.  The flags are
	// the same few words for the whole grid and same-address atomics serialise in L2 (one set per warp on ONE address cost
	// ~100 us per million sources): a flag is only OR-ed by the warps that started while it still read 0, and the counters go to one of kStatParts
	// padded slots chosen by the block index (K4b sums them).  No block-wide barrier: a finished warp leaves the SM.
	const unsigned full = 0xffffffffu;
	uint32_t u = used;
	unsigned long long bytes = n_src ? 4ull * n_src : 0ull;
	uint32_t rendered = n_src ? 1u : 0u;
	for (int o = 16; o > 0; o >>= 1) {
		u |= __shfl_xor_sync(full, u, o);
		bytes += __shfl_xor_sync(full, bytes, o);
		rendered += __shfl_xor_sync(full, rendered, o);
	}
	if ((threadIdx.x & 31) == 0) {
		if ((u & kUsedPan) && !set_pan) atomicOr(&flags->pan_used[g.parity], 1);
		if ((u & kUsedSh) && !set_sh) atomicOr(&flags->sh_used[g.parity], 1);
		if ((u & kUsedStereo) && !set_st) atomicOr(&flags->st_used[g.parity], 1);
		if ((u & kUsedRefl) && !set_refl) atomicOr(&flags->refl_used[g.parity], 1);
		if (rendered) {
			const uint32_t part = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % kStatParts;
			atomicAdd(&flags->part_bytes[part][0], bytes);
			atomicAdd(&flags->part_rendered[part][0], rendered);
		}
	}
}

__global__ void __launch_bounds__(256) k_patch_kernel(SlotArrays A, const SlotPatch* __restrict__ patches, uint32_t n,
                                                     const MixPatch* __restrict__ mp, uint32_t nmp) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		const SlotPatch p = patches[i];
		const uint32_t s = p.slot;
		if (p.what & kPatchReset) {
			A.live[s] = 1; A.read_i[s] = 0; A.eof[s] = p.eof;
			A.dgain[s] = 0.0f; A.occ_cur[s] = 0.0f; A.lpf_prev[s] = 0.0f; A.bits[s] = 0;
		}
		if (p.what & kPatchReadI) A.read_i[s] = p.read_i;
		if (p.what & kPatchEof) A.eof[s] = p.eof;
		if (p.what & (kPatchResampler | kPatchReset)) {
			// AudioResampler::init (AR.cpp:24-36)
			Resamp r; r.prev_dest_dst_coords = -1; r.prev_samples_0_src_coords = -2; r.prev_samples[0] = 0.0f; r.prev_samples[1] = 0.0f;
			A.rs[s] = r;
		}
		if (p.what & kPatchKill) A.live[s] = 0;
	}
	if (i < nmp) { A.mix_pos[mp[i].index] = mp[i].pos; A.mix_pos0[mp[i].index] = mp[i].pos; }
}

namespace dev {

int launch_patch(const SlotArrays& A, const SlotPatch* patches, uint32_t n, const MixPatch* mp, uint32_t nmp, void* stream) {
	const uint32_t m = n > nmp ? n : nmp;
	if (!m) return 0;
	k_patch_kernel<<<(m + 255) / 256, 256, 0, (cudaStream_t)stream>>>(A, patches, n, mp, nmp);
	note_launch();
	return check_launch("k_patch");
}

int launch_params(const Globals& g, const SlotArrays& A, DevFlags* flags, const float* table, const uint32_t* lists, uint32_t n_slots, void* stream) {
	(void)n_slots;
	const uint32_t n = g.n_pan + g.n_sh;
	if (!n) return 0;
	k1_params_kernel<<<(n + SBM_K1_THREADS - 1) / SBM_K1_THREADS, SBM_K1_THREADS, 0, (cudaStream_t)stream>>>(g, A, flags, table, lists, window_slot(g.frames));
	note_launch();
	return check_launch("k1_params");
}

}  // namespace dev
}  // namespace sbm
