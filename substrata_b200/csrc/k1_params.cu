// k1_params.cu -- "K1": one thread per render-list entry (live source) runs the feeder bookkeeping and Resonance's per-source parameter
// logic for this buffer (k1_core.h: update_source) and condenses the slot into one WorkRec for the mix kernel; plus
// k_patch, which applies host-requested state (re)initialisations before K1.
//
// Compiled with -fmad=false: results of this unit feed integer truncations and thresholds (lookup-table cell,
// -60 dB rules, ramp lengths), so every float operation must round exactly like the reference's SSE build.
// HBM traffic per slot: SrcStatic 32 B + pos/volume 16 B + ~60 B of state read-modify-write + 128 B MixRec write (k2_rec.h;
// + the 128 B WorkRec only for sources a careful path of the mix kernel will read back).
#include <atomic>

#include "device_internal.h"
#include "k1_core.h"
#include "k2_rec.h"

namespace sbm {

#ifndef SBM_K1_THREADS
#define SBM_K1_THREADS 128             // 64 .. 256 measure within 3 %; 512 is 5 % slower (one thread per source, latency-bound)
#endif
#ifndef SBM_K1_MINB
#define SBM_K1_MINB 6                  // 80 registers, no spills; K1 runs the same at 768 and 1024 resident threads (it is bound by its memory transactions)
#endif
// Overlapped mode (Globals::overlap; large stereo-bus scenes): the grid is PERSISTENT -- kK1OverlapCtas CTAs per SM, sized so that the
// mix kernel's CTAs fit beside them -- and walks the render list tile by tile in the order the mix kernel consumes it (that kernel deals
// batches to its CTAs round-robin, so the whole grid sweeps the list front to back).  Every CTA announces the dependent launch at once; the
// mix kernel (k2_mix.cu: wait_records_ready) then polls SlotArrays::ready before it fetches a tile's records.  K1 never waits for K2 and is
// fully resident before K2's first CTA can start, so the pair cannot deadlock.  K1 alone is latency-bound (3 dependent gathers per source)
// and K2 is HBM-bound: side by side, K1's time disappears under K2's, and K2 finds the records K1 just wrote in L2.
constexpr int kK1OverlapCtas = 2;
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
static_assert(SBM_K1_THREADS == (int)kK1Tile, "one thread per entry of a K1 tile");

// Record stores.  One thread's 128-byte MixRec (or the 64 + 192 bytes of ShGains / ShRamps of an ambisonic source) written by that thread alone
// is 8 (4 + 12) stores that each touch 32 different lines -- K1 was bound by those store transactions, not by its loads (it ran
// the same at 768 and 1024 resident threads).  The records go to shared memory first and the WARP writes them out as consecutive
// 16-byte stores (a warp's 32 list entries are 4 KB / 2 KB / 6 KB of contiguous global memory): 1 M panned sources 0.100 -> 0.083 ms.
//   SH (the launch has ambisonic sources): a thread's staging slot is 272 bytes -- ShGains at 0, ShRamps at 64, padded so that the
//   threads' 16-byte stores spread over the banks -- written by update_source itself, copied out, then reused for the MixRec.
template <bool OVERLAP, bool SH>
__global__ void __launch_bounds__(SBM_K1_THREADS, OVERLAP ? 8 : SBM_K1_MINB) /* beside K2: 64 registers, or K2's CTAs do not fit */ k1_params_kernel(Globals g, SlotArrays A, DevFlags* flags, const float* __restrict__ table,
                                                       const uint32_t* __restrict__ lists, uint32_t wslot) {
	constexpr uint32_t kSlotBytes = SH ? 272u : 128u;
	__shared__ __align__(16) unsigned char stage_raw[SBM_K1_THREADS * kSlotBytes];
	unsigned char* wstage = stage_raw + (threadIdx.x >> 5) * 32u * kSlotBytes;      // this warp's staging region
	const uint32_t lane = threadIdx.x & 31u;
	const unsigned full = 0xffffffffu;
	if (OVERLAP) grid_dep_launch_dependents();
	const uint32_t n_list = g.n_pan + g.n_sh, n_tiles = (n_list + kK1Tile - 1) / kK1Tile;
	// whether this buffer's "mixer used" flags are already set: read NOW, so that the loads overlap the work below instead
	// of being a dependent round trip at the end of every warp (a stale 0 only costs a redundant atomicOr)
	int set_pan = flags->pan_used[g.parity], set_sh = flags->sh_used[g.parity], set_st = flags->st_used[g.parity], set_refl = flags->refl_used[g.parity];
	// Persistent grid: few threads per SM, so the chain  list entry -> slot -> a dozen per-slot arrays  is taken off the critical path by
	// software: the list entry is read two tiles ahead and the slot's lines are prefetched into L2 one tile ahead.
	const uint32_t stride = OVERLAP ? gridDim.x * kK1Tile : 0u;
	// Classic mode walks the list BACK TO FRONT (CTAs are dispatched in index order): the records written last are the ones the mix kernel
	// reads first, so the head of the list is still in L2 when it starts.
	const uint32_t tile0 = OVERLAP ? blockIdx.x : n_tiles - 1u - blockIdx.x;
	uint32_t k0 = tile0 * kK1Tile + threadIdx.x;
	const bool ident = g.list_identity != 0;
	uint32_t slot_cur = k0 < n_list ? (ident ? k0 + g.list_base : lists[k0]) : 0u, slot_n1 = 0u;
	if (OVERLAP && k0 + stride < n_list) slot_n1 = ident ? k0 + stride + g.list_base : lists[k0 + stride];
	for (uint32_t tile = tile0; tile < n_tiles; tile += OVERLAP ? gridDim.x : n_tiles) {      // (classic: one tile per CTA)
		const uint32_t k = tile * kK1Tile + threadIdx.x;    // position in the render lists
		const uint32_t kw = tile * kK1Tile + (threadIdx.x & ~31u);     // first list entry of this warp
		uint32_t used = 0, n_src = 0;
		uint32_t slot_n2 = 0u;
		if (OVERLAP) {
			if (k + 2u * stride < n_list) slot_n2 = ident ? k + 2u * stride + g.list_base : lists[k + 2u * stride];
			if (k + stride < n_list) {
				const uint32_t sn = slot_n1;
				prefetch_l2(A.stat + sn); prefetch_l2(A.posvol + sn); prefetch_l2(A.occl + sn); prefetch_l2(A.live + sn); prefetch_l2(A.read_i + sn);
				prefetch_l2(A.eof + sn); prefetch_l2(A.rs + sn); prefetch_l2(A.dgain + sn); prefetch_l2(A.occ_cur + sn); prefetch_l2(A.lpf_prev + sn);
				prefetch_l2(A.bits + sn); prefetch_l2(A.cgain + (size_t)sn * 4u);
			}
		}
		const bool valid = k < n_list;
		MixRec m;
		bool sh_ramps = false;
		if (valid) {
			WorkRec rec;
			ReflSend rfl;
			rfl.konst = 0.0f; rfl.start = 0.0f; rfl.inc = 0.0f; rfl.n = 0;
			const uint32_t mix_hint = g.mix_identity ? k + g.mix_base : (A.list_mix ? A.list_mix[k] : 0xffffffffu);
			ShGains* shg_dst = SH ? reinterpret_cast<ShGains*>(wstage + lane * kSlotBytes) : A.shg + k;
			ShRamps* shr_dst = SH ? reinterpret_cast<ShRamps*>(wstage + lane * kSlotBytes + 64u) : A.shr + k;
			used = update_source(g, A, flags, table, slot_cur, mix_hint, &rec, shg_dst, shr_dst, &rfl, &n_src);
			sh_ramps = (rec.flags & kRecHrtf) && rec.ch_ramp_mask != 0;
			// the record the mix kernel consumes, in final form (k2_rec.h); the full WorkRec only where a careful path will read it back
			m = make_mixrec(rec, wslot, k >= g.n_pan);
			if (g.room) {
				// a ramping reflections send needs the per-frame gain path like any other ramp
				A.rfl[k] = rfl;
				if (rfl.n && !(m.mode & kMdSkip)) { if (rfl.n > m.nramp) m.nramp = rfl.n; m.mode |= kMdRamp; }
			}
			if (mixrec_needs_workrec(m.mode)) A.rec[k] = rec;
		}
		if (SH) {
			// the warp's ShGains (4 chunks of 16 bytes per entry) and, for the entries with channel ramps, ShRamps (12 chunks)
			__syncwarp();
			const unsigned m_sh = __ballot_sync(full, valid && k >= g.n_pan), m_rp = __ballot_sync(full, sh_ramps);
			if (m_sh) {
				uint4* gg = reinterpret_cast<uint4*>(A.shg + kw);
#pragma unroll
				for (uint32_t i = 0; i < 4; ++i) {
					const uint32_t idx = i * 32u + lane, r = idx >> 2, c = idx & 3u;
					if ((m_sh >> r) & 1u) gg[idx] = *reinterpret_cast<const uint4*>(wstage + r * kSlotBytes + c * 16u);
				}
			}
			if (m_rp) {
				uint4* gr = reinterpret_cast<uint4*>(A.shr + kw);
#pragma unroll
				for (uint32_t i = 0; i < 12; ++i) {
					const uint32_t idx = i * 32u + lane, r = idx / 12u, c = idx - r * 12u;
					if ((m_rp >> r) & 1u) gr[idx] = *reinterpret_cast<const uint4*>(wstage + r * kSlotBytes + 64u + c * 16u);
				}
			}
			__syncwarp();      // the region is reused for the MixRecs
		}
		{
			// MixRec: 16-byte chunks swizzled by lane on the way in (4 wavefronts per store instead of 32), consecutive on the way out
			uint4* my = reinterpret_cast<uint4*>(wstage);
			if (valid) {
				uint4 mv[8];
				__builtin_memcpy(mv, &m, sizeof(MixRec));
#pragma unroll
				for (uint32_t j = 0; j < 8; ++j) my[lane * 8u + (j ^ (lane & 7u))] = mv[j];
			}
			__syncwarp();
			uint4* gw = reinterpret_cast<uint4*>(A.mrec + kw);
#pragma unroll
			for (uint32_t i = 0; i < 8; ++i) {
				const uint32_t idx = i * 32u + lane, r = idx >> 3, c = idx & 7u;
				if (kw + r < n_list) gw[idx] = my[r * 8u + (c ^ (r & 7u))];
			}
			if (OVERLAP || SH) __syncwarp();      // (the stage is reused: by the next tile of a persistent grid)
		}
		// Warp-level reduction of the "mixer used" bits and the accounting counters, then one lane publishes.  The flags are
		// the same few words for the whole grid and same-address atomics serialise in L2 (one set per warp on ONE address cost
		// ~100 us per million sources): a flag is only OR-ed by the warps that started while it still read 0, and the counters go to one of kStatParts
		// padded slots chosen by the block index (K4b sums them).  No block-wide barrier (classic mode): a finished warp leaves the SM.
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
				const uint32_t part = (tile * (blockDim.x >> 5) + (threadIdx.x >> 5)) % kStatParts;
				atomicAdd(&flags->part_bytes[part][0], bytes);
				atomicAdd(&flags->part_rendered[part][0], rendered);
			}
		}
		// (a persistent warp sets each flag at most once)
		if (OVERLAP) {
			slot_cur = slot_n1; slot_n1 = slot_n2;
			set_pan |= (int)(u & kUsedPan); set_sh |= (int)(u & kUsedSh); set_st |= (int)(u & kUsedStereo); set_refl |= (int)(u & kUsedRefl);
			// publish the tile: every thread's stores (records, state, counters) ordered before the stamp -- also for the mix kernel's
			// bulk copies, which read through the async proxy
			__threadfence();
			asm volatile("fence.proxy.async.global;" ::: "memory");
			__syncthreads();
			if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(A.ready + tile), "r"(g.epoch) : "memory");
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
	uint32_t grid = (n + SBM_K1_THREADS - 1) / SBM_K1_THREADS;
	if (g.overlap) {
		const uint32_t resident = (uint32_t)(current_sm_count() * kK1OverlapCtas);
		if (grid > resident) grid = resident;
	}
	if (g.overlap) {
		// an SM holds ONE shared-memory carve-out at a time: K1 must ask for the mix kernel's (maximum shared memory), or that kernel's
		// CTAs could not join K1's on the same SM before K1 has drained
		static std::atomic<unsigned long long> configured{0};
		int device = 0;
		cudaGetDevice(&device);
		const unsigned long long bit = 1ull << (device & 63);
		if (!(configured.load(std::memory_order_acquire) & bit)) {
			if (check(cudaFuncSetAttribute(k1_params_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), "cudaFuncSetAttribute(k1 carveout)")) return -1;
			configured.fetch_or(bit, std::memory_order_release);
		}
	}
	if (g.overlap) k1_params_kernel<true, false><<<grid, SBM_K1_THREADS, 0, (cudaStream_t)stream>>>(g, A, flags, table, lists, window_slot(g.frames));
	else if (g.n_sh) k1_params_kernel<false, true><<<grid, SBM_K1_THREADS, 0, (cudaStream_t)stream>>>(g, A, flags, table, lists, window_slot(g.frames));
	else k1_params_kernel<false, false><<<grid, SBM_K1_THREADS, 0, (cudaStream_t)stream>>>(g, A, flags, table, lists, window_slot(g.frames));
	note_launch();
	return check_launch("k1_params");
}

}  // namespace dev
}  // namespace sbm
