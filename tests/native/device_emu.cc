// tests/native/device_emu.cc -- TEST INFRASTRUCTURE ONLY.  Never compiled into libsubstrata_b200.so.
//
// Host emulation of the device seam (substrata_b200/csrc/device_api.h) so that the no-GPU test suite can drive the
// REAL host engine (engine.cc: id/slot tables, dirty-range uploads, patch merging, ring bookkeeping, finished-source
// handling) and the REAL per-source / per-sample arithmetic (k1_core.h, k2_core.h -- the very functions the CUDA
// kernels call) against the oracle on CPU.  What it does not cover is the kernels' parallel decomposition, which only
// exists on the GPU and is checked there (`pytest -m gpu`).  Linked into tests/native/libsbm_emu.so only.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "device_api.h"
#include "k1_core.h"
#include "k2_core.h"

namespace sbm {
namespace dev {

static unsigned long long g_launches = 0;

int init(int, void** stream_out) { *stream_out = (void*)1; return 0; }
void* stream_create() { return (void*)2; }
int use(int) { return 0; }
void shutdown(void*) {}
const char* last_error() { return "emulator"; }
void* alloc(size_t bytes) { return calloc(bytes ? bytes : 1, 1); }
void release(void* p) { free(p); }
void* alloc_pinned(size_t bytes) { return calloc(bytes ? bytes : 1, 1); }
void release_pinned(void* p) { free(p); }
bool is_pinned(const void*) { return false; }
int h2d(void* dst, const void* src, size_t bytes, void*) { memcpy(dst, src, bytes); return 0; }
int d2h(void* dst, const void* src, size_t bytes, void*) { memcpy(dst, src, bytes); return 0; }
int zero(void* dst, size_t bytes, void*) { memset(dst, 0, bytes); return 0; }
int sync(void*) { return 0; }
int num_mix_ctas(uint32_t, int) { return 1; }
size_t partials_floats(uint32_t) { return 1; }

int launch_apply_uploads(const UpSeg* segs, uint32_t n_segs, const uint32_t* block, uint32_t, void*) {
	++g_launches;
	for (uint32_t k = 0; k < n_segs; ++k) {
		const UpSeg& sg = segs[k];
		uint32_t* dst = (uint32_t*)sg.dst;
		const uint32_t* val = block + sg.src_word;
		for (uint32_t j = 0; j < sg.count; ++j) {
			const uint32_t at = sg.idx_word == 0xffffffffu ? j : block[sg.idx_word + j];
			memcpy(dst + (size_t)at * sg.elem_words, val + (size_t)j * sg.elem_words, 4 * (size_t)sg.elem_words);
		}
	}
	return 0;
}

int launch_patch(const SlotArrays& A, const SlotPatch* patches, uint32_t n, const MixPatch* mp, uint32_t nmp, void*) {
	++g_launches;
	for (uint32_t i = 0; i < n; ++i) {
		const SlotPatch p = patches[i];
		const uint32_t s = p.slot;
		if (p.what & kPatchReset) {
			A.live[s] = 1; A.read_i[s] = 0; A.eof[s] = p.eof;
			A.dgain[s] = 0.0f; A.occ_cur[s] = 0.0f; A.lpf_prev[s] = 0.0f; A.bits[s] = 0;
		}
		if (p.what & kPatchReadI) A.read_i[s] = p.read_i;
		if (p.what & kPatchEof) A.eof[s] = p.eof;
		if (p.what & (kPatchResampler | kPatchReset)) {
			Resamp r; r.prev_dest_dst_coords = -1; r.prev_samples_0_src_coords = -2; r.prev_samples[0] = 0.0f; r.prev_samples[1] = 0.0f;
			A.rs[s] = r;
		}
		if (p.what & kPatchKill) A.live[s] = 0;
	}
	for (uint32_t i = 0; i < nmp; ++i) { A.mix_pos[mp[i].index] = mp[i].pos; A.mix_pos0[mp[i].index] = mp[i].pos; }
	return 0;
}

int launch_params(const Globals& g, const SlotArrays& A, DevFlags* flags, const float* table, const uint32_t* lists, uint32_t n_slots, void*) {
	++g_launches;
	uint32_t used = 0;
	(void)n_slots;
	for (uint32_t k = 0; k < g.n_pan + g.n_sh; ++k) {
		uint32_t n_src = 0;
		ReflSend rfl; rfl.konst = 0.0f; rfl.start = 0.0f; rfl.inc = 0.0f; rfl.n = 0;
		const uint32_t slot = g.list_identity ? k + g.list_base : lists[k];
		used |= update_source(g, A, flags, table, slot, g.mix_identity ? k + g.mix_base : (A.list_mix ? A.list_mix[k] : 0xffffffffu), &A.rec[k], A.shg ? A.shg + k : nullptr, A.shr ? A.shr + k : nullptr, &rfl, &n_src);
		if (g.room) A.rfl[k] = rfl;
		if (n_src) { flags->algo_bytes += 4ull * n_src; flags->n_rendered++; }
	}
	if (used & kUsedRefl) flags->refl_used[g.parity] = 1;
	if (used & kUsedPan) flags->pan_used[g.parity] = 1;
	if (used & kUsedSh) flags->sh_used[g.parity] = 1;
	if (used & kUsedStereo) flags->st_used[g.parity] = 1;
	return 0;
}

static void mix_list(const MixArgs& a, uint32_t first, uint32_t n, bool sh) {
	const uint32_t F = a.frames;
	std::vector<float> y(F);
	for (uint32_t k = first; k < first + n; ++k) {
		const WorkRec& r = a.rec[k];
		if (r.flags & kRecSkip) {
			if (r.flags & kRecResample) resampler_tail(r, &a.rs[r.slot].prev_samples[0], &a.rs[r.slot].prev_samples[1]);
			continue;
		}
		if (a.rfl && !(r.flags & kRecStereo)) {
			// room effects: the raw signal through the ramped reflections send (graph_manager.cc:150)
			RampDesc d; d.konst = a.rfl[k].konst; d.start = a.rfl[k].start; d.inc = a.rfl[k].inc; d.n = a.rfl[k].n;
			float* row = a.bus + (size_t)kRowRefl * F;
			for (uint32_t i = 0; i < F; ++i) row[i] = fmaf(gain_at(d, i), source_sample(r, a.mixz, a.mix_pos0, i), row[i]);
		}
		for (uint32_t i = 0; i < F; ++i) y[i] = fmul(source_sample(r, a.mixz, a.mix_pos0, i), gain_at(r.direct, i));
		if (r.flags & kRecResample) resampler_tail(r, &a.rs[r.slot].prev_samples[0], &a.rs[r.slot].prev_samples[1]);
		if (!(r.flags & kRecStereo)) {
			if (r.flags & kRecLpf) {
				float p = r.lpf_prev;
				for (uint32_t i = 0; i < F; ++i) { p = lpf_step(r.lpf_coef, p, y[i]); y[i] = p; }
				a.lpf_prev[r.slot] = p;
			} else {
				a.lpf_prev[r.slot] = y[F - 1];
			}
		}
		if (!sh) {
			for (int c = 0; c < 2; ++c) {
				float* row = a.bus + (size_t)(16 + c) * F;
				for (uint32_t i = 0; i < F; ++i) row[i] = fmaf(gain_at(r.ch[c], i), y[i], row[i]);
			}
		} else {
			for (int c = 0; c < kNumSh; ++c) {
				const float konst = a.shg[k].konst[c];
				const bool ramps = (r.ch_ramp_mask & (1u << c)) != 0;
				float* row = a.bus + (size_t)c * F;
				for (uint32_t i = 0; i < F; ++i) row[i] = fmaf(sh_gain_at(konst, ramps ? &a.shr[k] : nullptr, c, i), y[i], row[i]);
			}
		}
	}
}

int launch_mix(const MixArgs& a, void*) {
	g_launches += 2;
	const uint32_t F = a.frames;
	memset(a.bus, 0, sizeof(float) * (size_t)kBusRows * F);
	mix_list(a, 0, a.n_pan, false);
	mix_list(a, a.n_pan, a.n_sh, true);
	float* bf = a.bus + (size_t)kBusRows * F;
	bf[0] = a.flags->sh_used[a.parity] ? 1.0f : 0.0f;
	bf[1] = a.flags->pan_used[a.parity] ? 1.0f : 0.0f;
	bf[2] = a.flags->st_used[a.parity] ? 1.0f : 0.0f;
	bf[3] = a.flags->refl_used[a.parity] ? 1.0f : 0.0f;
	return 0;
}

// The host emulation has no peer GPUs: multi-rank CPU tests sum the buses with a gloo all-reduce instead (sharding.py).
int peer_export(void*, void*) { return -1; }
void* peer_open(const void*) { return nullptr; }
void peer_close(void*) {}
int launch_peer_gather(const PeerArgs&, const float*, float*, size_t, unsigned int, DevFlags*, void*) { return -1; }

int launch_ring_scatter(const RingPush* packets, uint32_t n, const float* stage, void*) {
	if (!n) return 0;
	++g_launches;
	for (uint32_t k = 0; k < n; ++k) {
		const RingPush& p = packets[k];
		for (uint32_t i = 0; i < p.n; ++i) p.ring[(p.at + i) % p.capacity] = stage[p.src_off + i];
	}
	return 0;
}

int launch_polyphase(const PolyphaseArgs& a, void*) {
	++g_launches;
	for (uint32_t c = 0; c < a.channels; ++c)
		for (uint32_t k = 0; k < a.n_out; ++k) {
			const unsigned long long total = (unsigned long long)a.time_modulo + (unsigned long long)k * a.down;
			const uint32_t input_sample = a.last + (uint32_t)(total / a.up), phase = (uint32_t)(total % a.up);
			const float* x = a.in + (size_t)c * a.ext + (size_t)input_sample;
			const float* h = a.coeffs + (size_t)phase * a.cpp;
			float acc = 0.0f;
			for (uint32_t j = 0; j < a.cpp; ++j) acc = fadd(acc, fmul(x[j], h[j]));
			a.out[(size_t)c * a.n_out + k] = acc;
		}
	return 0;
}

int launch_room(const RoomArgs& a, void*) {
	++g_launches;
	const uint32_t F = a.frames;
	const float* in = a.bus + (size_t)kRowRefl * F;
	RoomPlan plan;
	const RoomParams p = *a.params;
	room_plan(p, a.state, p.enabled && a.bus[(size_t)kBusRows * F + 3] > 0.0f, F, &plan);
	a.flags->foa_valid = plan.process ? 1 : 0;
	if (!plan.process) return 0;
	std::vector<float> lp(F);
	for (uint32_t i = 0; i < F; ++i) lp[i] = plan.silent_input ? 0.0f : in[i];
	if (!(plan.lp_coef < kEpsilonFloat)) a.state->lp_prev = room_lowpass(plan.lp_coef, plan.lp_prev, lp.data(), lp.data(), F);
	else a.state->lp_prev = lp[F - 1];
	for (uint32_t i = 0; i < F; ++i) a.delay_line[room_insert_index(plan, F, i)] = lp[i];
	for (uint32_t i = 0; i < F; ++i) room_frame(plan, a.delay_line, F, i, &a.foa[i], &a.foa[(size_t)F + i], &a.foa[(size_t)2 * F + i], &a.foa[(size_t)3 * F + i]);
	return 0;
}

int launch_finish(const FinishArgs& a, void*) {
	++g_launches;
	const uint32_t F = a.frames;
	const int foa_valid = a.hrir1 ? a.flags->foa_valid : 0;
	const FinishPlan plan = plan_finish(a.bus + (size_t)kBusRows * F, a.hrir != nullptr, a.hrir1 != nullptr, a.room_enabled != 0, a.flags, a.parity, F);
	std::vector<float> l(F, 0.0f), r(F, 0.0f);
	if (plan.decode1) {
		// the reflections' order-1 decode (silence in while the decoder only flushes its tail)
		std::vector<float> foa(4 * (size_t)F, 0.0f);
		if (foa_valid) memcpy(foa.data(), a.foa, sizeof(float) * 4 * F);
		const int H1 = (int)a.taps1 - 1;
		for (uint32_t i = 0; i < F; ++i) decode_frame(foa.data(), a.hrir1, a.hist1, F, a.taps1, i, &l[i], &r[i], 4);
		if (H1 > 0) {
			std::vector<float> nh((size_t)H1);
			for (int c = 0; c < 4; ++c) {
				float* hc = a.hist1 + (size_t)c * (size_t)H1;
				for (int q = 0; q < H1; ++q) nh[(size_t)q] = decode_input(foa.data() + (size_t)c * F, hc, H1, (int)F - H1 + q);
				memcpy(hc, nh.data(), sizeof(float) * (size_t)H1);
			}
		}
	}
	std::vector<float> l3(F, 0.0f), r3(F, 0.0f);
	if (plan.decode) {
		const int H = (int)a.taps - 1;
		for (uint32_t i = 0; i < F; ++i) { decode_frame(a.bus, a.hrir, a.hist, F, a.taps, i, &l3[i], &r3[i]); l[i] += l3[i]; r[i] += r3[i]; }
		if (H > 0) {
			std::vector<float> nh((size_t)H);
			for (int c = 0; c < kNumSh; ++c) {
				float* hc = a.hist + (size_t)c * (size_t)H;
				for (int q = 0; q < H; ++q) nh[(size_t)q] = decode_input(a.bus + (size_t)c * F, hc, H, (int)F - H + q);
				memcpy(hc, nh.data(), sizeof(float) * (size_t)H);
			}
		}
	}
	for (uint32_t i = 0; i < F; ++i) {
		float lo = l[i], ro = r[i];
		if (plan.stereo) { lo += a.bus[(size_t)16 * F + i]; ro += a.bus[(size_t)17 * F + i]; }
		if (!plan.valid) { lo = 0.0f; ro = 0.0f; }
		a.out[2 * i] = lo; a.out[2 * i + 1] = ro;
	}
	OutTail* tail = reinterpret_cast<OutTail*>(a.out + 2 * F);
	tail->valid = plan.valid; tail->n_finished = a.flags->n_finished; tail->n_rendered = a.flags->n_rendered; tail->status = kTailOk;
	tail->algo_bytes = a.flags->algo_bytes;
	a.flags->n_finished = 0; a.flags->n_rendered = 0; a.flags->algo_bytes = 0;
	unsigned long long* reads = reinterpret_cast<unsigned long long*>(tail + 1);
	for (uint32_t q = 0; q < a.n_ring; ++q) reads[q] = a.read_i[a.list_ring[q]];
	int* fin = reinterpret_cast<int*>(reads + a.n_ring);
	for (int q = 0; q < tail->n_finished && q < kInlineFinished; ++q) fin[q] = a.finished[q];
	return 0;
}

void* event_create() { return nullptr; }
void event_destroy(void*) {}
void event_record(void*, void*) {}
int event_sync(void*) { return 0; }
int stream_wait_event(void*, void*) { return 0; }
float event_elapsed_ms(void*, void*) { return 0.0f; }
unsigned long long launches() { return g_launches; }

}  // namespace dev
}  // namespace sbm
