// engine.cc -- host side of the B200 mixer behind the C ABI of include/sbm_audio.h.
//
// Owns: the id -> slot table, host mirrors of the host-written per-slot parameters (with dirty ranges so that one
// sbm_mix uploads only what the setters touched), the PCM buffer registry, ring storage for streamed sources, and
// the per-buffer launch sequence  patch -> K1 (params) -> K2 (mix) -> K3 (reduce) -> [all-reduce by the caller] ->
// K4 (decode + stereo mix + interleave) -> D2H.  All arithmetic on audio happens in the kernels (device_api.h);
// nothing here touches a sample except to copy pushed ring packets to the device.
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sbm_audio.h"
#include "device_api.h"

namespace sbm {
const float* encoder_table();   // sh_table.cc
}

using namespace sbm;

namespace {

struct DirtyRange {
	uint32_t lo = 0xffffffffu, hi = 0;
	void mark(uint32_t i) { if (i < lo) lo = i; if (i + 1 > hi) hi = i + 1; }
	bool any() const { return hi > lo; }
	void clear() { lo = 0xffffffffu; hi = 0; }
};

struct BufferInfo { float* dptr; size_t n; bool owned; };

struct RingInfo { float* storage; uint32_t capacity; unsigned long long written; unsigned long long read_known; };

}  // namespace

struct sbm_engine {
	sbm_config cfg;
	void* stream = nullptr;
	std::string error;
	std::mutex mu;               // guards everything the setters touch

	uint32_t cap = 0;            // slot capacity
	uint32_t n_slots = 0;        // high-water mark
	std::vector<uint32_t> free_slots;
	std::vector<int> slot_of_id; // public id -> slot (-1 = dead)
	int next_id = 0;

	// host mirrors of host-written arrays
	std::vector<SrcStatic> h_stat;
	std::vector<F4> h_posvol;
	std::vector<float> h_occl;
	std::vector<unsigned long long> h_ring_write;
	std::vector<MixZ> h_mixz;
	DirtyRange d_stat, d_posvol, d_occl, d_ring, d_mixz;
	uint32_t mix_used = 0;       // MixZ entries handed out (bump allocation; freed entries are not recycled in this round)
	std::vector<SlotPatch> patches;          // at most one (merged) pending patch per slot
	std::vector<int> patch_of_slot;          // slot -> index into patches, -1 = none pending
	std::vector<MixPatch> mix_patches;
	std::vector<RingInfo> rings;            // indexed by slot (storage == nullptr when not a ring source)
	std::vector<uint32_t> list_pan, list_sh, list_ring;
	bool lists_dirty = true;
	std::vector<BufferInfo> buffers;
	std::vector<int> finished_queue;

	// device
	SlotArrays A;                // device pointers
	SrcStatic* dv_stat = nullptr; F4* dv_posvol = nullptr; float* dv_occl = nullptr; unsigned long long* dv_ring_write = nullptr;
	MixZ* dv_mixz = nullptr;
	DevFlags* dv_flags = nullptr;
	float* dv_table = nullptr;
	uint32_t *dv_lists = nullptr, *dv_list_ring = nullptr;   // dv_lists: pan list followed by the ambisonic list
	SlotPatch* dv_patches = nullptr; uint32_t dv_patches_cap = 0;
	MixPatch* dv_mix_patches = nullptr; uint32_t dv_mix_patches_cap = 0;
	float* dv_partials = nullptr; size_t partials_floats = 0;
	float* dv_bus = nullptr; size_t bus_floats = 0;
	float* dv_hrir = nullptr; float* dv_hist = nullptr; float* dv_dec = nullptr;
	float* dv_out = nullptr; size_t out_bytes = 0;
	// multi-GPU bus exchange over peer memory (sbm_peer_export / sbm_peer_connect)
	PeerArgs peer{};                         // world == 0 until connected
	float* dv_gather = nullptr;              // own gather area [2][kMaxPeers][bus_floats] followed by the counters
	float* dv_bus_sum = nullptr;             // bus summed over ranks (what K4 reads when connected)
	void* peer_mapped[kMaxPeers] = {};       // IPC mappings of the other ranks' areas
	unsigned char* h_ring_stage = nullptr;   // pinned staging of sbm_source_ring_push_batch: packet table + payload
	unsigned char* d_ring_stage = nullptr;
	size_t ring_stage_bytes = 0;
	unsigned long long peer_epoch = 0;       // buffer_index at which the peers were connected (arrival counters count from there)
	size_t pending_ring_h2d_bytes = 0;       // ring pushes since the last sbm_mix_begin (reported with the next buffer's stats)
	bool ring_stage_in_flight = false;       // an H2D from h_ring_stage may still be running (cleared by every stream sync in mix)
	// Up to two buffers in flight (sbm_mix_submit / sbm_mix_collect): everything a submitted buffer's completion needs.
	struct InFlight {
		bool pending = false;
		unsigned long long seq = 0;          // buffer_index of the buffer
		unsigned char* h_out = nullptr;      // pinned D2H staging: 2F floats + OutTail + ring read counters
		void* done = nullptr;                // event after the D2H
		void* ev[4] = {nullptr, nullptr, nullptr, nullptr};   // profiling: before K1 / after K1 / after K2 / after K4
		int* dv_finished = nullptr;          // this parity's finished-id list on the device
		std::vector<uint32_t> list_ring;     // ring-fed slots in the order their read counters were reported
		size_t h2d_bytes = 0, d2h_bytes = 0;
		uint32_t active = 0;
	} fl[2];
	unsigned long long collected = 0;        // buffers collected so far (== buffer_index when nothing is in flight)
	unsigned long long ring_stage_seq = 0;   // the staged ring push precedes buffer `ring_stage_seq` in stream order
	Globals g;
	unsigned long long buffer_index = 0;
	bool begun = false;
	int last_valid = 0;

	// stats
	bool profiling = false;
	sbm_stats stats;
};

namespace {

int fail(sbm_engine* e, int code, const std::string& msg) {
	if (e) e->error = msg;
	return code;
}
int fail_cuda(sbm_engine* e, const char* where) {
	return fail(e, SBM_ERR_CUDA, std::string(where) + ": " + dev::last_error());
}

template <class T> bool dalloc(T*& p, size_t count) {
	p = (T*)dev::alloc(sizeof(T) * (count ? count : 1));
	return p != nullptr;
}

int slot_for(sbm_engine* e, int id) {
	if (id < 0 || (size_t)id >= e->slot_of_id.size()) return -1;
	return e->slot_of_id[(size_t)id];
}

// Queues a state (re)initialisation for k_patch.  Patches of one slot are merged so that the kernel can apply the
// list in parallel: a Reset discards what was pending, later fields override earlier ones.
void add_patch(sbm_engine* e, const SlotPatch& p) {
	int& at = e->patch_of_slot[p.slot];
	if (at < 0 || (p.what & kPatchReset)) {
		if (at < 0) { at = (int)e->patches.size(); e->patches.push_back(p); }
		else e->patches[(size_t)at] = p;
		return;
	}
	SlotPatch& q = e->patches[(size_t)at];
	q.what |= p.what;
	if (p.what & kPatchReadI) q.read_i = p.read_i;
	if (p.what & kPatchEof) q.eof = p.eof;
	if (p.what & kPatchResampler) q.rate = p.rate;
}

void rebuild_lists(sbm_engine* e) {
	e->list_pan.clear(); e->list_sh.clear(); e->list_ring.clear();
	for (uint32_t s = 0; s < e->n_slots; ++s) {
		const SrcStatic& st = e->h_stat[s];
		if (!(st.flags & kFlagAlive)) continue;
		const uint32_t kind = (st.flags >> kFlagKindShift) & 0xff;
		if (kind == kKindHrtf) e->list_sh.push_back(s); else e->list_pan.push_back(s);
		if (((st.flags >> kFlagFeedShift) & 3) == kFeedRing) e->list_ring.push_back(s);
	}
	e->lists_dirty = false;
}

void kill_slot(sbm_engine* e, uint32_t slot) {
	SrcStatic& st = e->h_stat[slot];
	if (!(st.flags & kFlagAlive)) return;
	if (st.id < e->slot_of_id.size()) e->slot_of_id[st.id] = -1;
	st.flags &= ~kFlagAlive;
	e->d_stat.mark(slot);
	if (e->rings[slot].storage) { dev::release(e->rings[slot].storage); e->rings[slot] = RingInfo{nullptr, 0, 0, 0}; }
	e->free_slots.push_back(slot);
	e->lists_dirty = true;
}

template <class T> int upload_range(sbm_engine* e, T* dv, const std::vector<T>& h, DirtyRange& d) {
	if (!d.any()) return 0;
	const size_t bytes = sizeof(T) * (size_t)(d.hi - d.lo);
	if (dev::h2d(dv + d.lo, h.data() + d.lo, bytes, e->stream)) return -1;
	e->stats.last_h2d_bytes += bytes;
	d.clear();
	return 0;
}

template <class T> int upload_list(sbm_engine* e, T*& dv, uint32_t& cap, const std::vector<T>& h) {
	if (h.empty()) return 0;
	if (h.size() > cap) {
		if (dv) { dev::sync(e->stream); dev::release(dv); }
		cap = (uint32_t)std::max<size_t>(h.size() * 2, 1024);
		if (!dalloc(dv, cap)) return -1;
	}
	if (dev::h2d(dv, h.data(), sizeof(T) * h.size(), e->stream)) return -1;
	e->stats.last_h2d_bytes += sizeof(T) * h.size();
	return 0;
}

}  // namespace

extern "C" {

SBM_API int sbm_create(const sbm_config* cfg, sbm_engine** out_engine) {
	if (!cfg || !out_engine) return SBM_ERR_INVALID_ARG;
	*out_engine = nullptr;
	// Resonance accepts 32..16384 frames (resonance_audio_api_impl.cc:78-94); this build tiles up to 1024.
	if (cfg->frames_per_buffer < 32 || cfg->frames_per_buffer > kMaxFrames) return SBM_ERR_INVALID_ARG;
	if (cfg->sample_rate_hz <= 0 || cfg->max_sources <= 0) return SBM_ERR_INVALID_ARG;
	if (cfg->sh_hrir && (cfg->sh_hrir_taps < 1 || cfg->sh_hrir_taps > 1024)) return SBM_ERR_INVALID_ARG;
	sbm_engine* e = new sbm_engine();
	e->cfg = *cfg;
	e->cfg.sh_hrir = nullptr;   // caller's pointer is not retained
	memset(&e->stats, 0, sizeof(e->stats));
	if (dev::init(cfg->device, &e->stream)) { const int rc = SBM_ERR_CUDA; fprintf(stderr, "sbm_create: %s\n", dev::last_error()); delete e; return rc; }
	const uint32_t cap = (uint32_t)cfg->max_sources;
	const uint32_t mixcap = (uint32_t)std::max(cfg->max_mix_sources, 1);
	const uint32_t F = (uint32_t)cfg->frames_per_buffer;
	e->cap = cap;
	e->h_stat.assign(cap, SrcStatic{nullptr, 0, 0, 0, 0, 0, 0});
	e->h_posvol.assign(cap, F4{0, 0, 0, 1});
	e->h_occl.assign(cap, 0.0f);
	e->h_ring_write.assign(cap, 0ull);
	e->h_mixz.assign(mixcap, MixZ{nullptr, 0, 0.0f, 0.0, 0.0});
	e->rings.assign(cap, RingInfo{nullptr, 0, 0, 0});
	e->patch_of_slot.assign(cap, -1);
	bool ok = true;
	SlotArrays& A = e->A;
	ok = ok && dalloc(e->dv_stat, cap) && dalloc(e->dv_posvol, cap) && dalloc(e->dv_occl, cap) && dalloc(e->dv_ring_write, cap) && dalloc(e->dv_mixz, mixcap);
	A.stat = e->dv_stat; A.posvol = e->dv_posvol; A.occl = e->dv_occl; A.ring_write = e->dv_ring_write; A.mixz = e->dv_mixz;
	ok = ok && dalloc(A.live, cap) && dalloc(A.read_i, cap) && dalloc(A.eof, cap) && dalloc(A.rs, cap) && dalloc(A.dgain, cap) &&
	     dalloc(A.occ_cur, cap) && dalloc(A.lpf_prev, cap) && dalloc(A.bits, cap) && dalloc(A.rec, cap) && dalloc(A.finished, 2 * (size_t)cap) &&
	     dalloc(A.mix_pos, mixcap) && dalloc(A.mix_pos0, mixcap);
	A.cstride = kNumSh;
	ok = ok && dalloc(A.cgain, (size_t)cap * kNumSh);
	const bool hrtf = cfg->sh_hrir != nullptr;
	if (hrtf) ok = ok && dalloc(A.shg, cap) && dalloc(A.shr, cap); else { A.shg = nullptr; A.shr = nullptr; }
	ok = ok && dalloc(e->dv_flags, 1) && dalloc(e->dv_table, (size_t)kTableAz * kTableEl * kShRaw);
	ok = ok && dalloc(e->dv_lists, cap) && dalloc(e->dv_list_ring, cap);
	e->bus_floats = (size_t)kBusRows * F + 4;
	ok = ok && dalloc(e->dv_bus, e->bus_floats);
	e->partials_floats = dev::partials_floats(F);
	ok = ok && dalloc(e->dv_partials, e->partials_floats);
	e->out_bytes = sizeof(float) * 2 * F + sizeof(OutTail) + sizeof(unsigned long long) * (size_t)cap;
	e->dv_out = (float*)dev::alloc(e->out_bytes);
	for (int k = 0; k < 2; ++k) {
		e->fl[k].h_out = (unsigned char*)dev::alloc_pinned(e->out_bytes);
		e->fl[k].dv_finished = A.finished ? A.finished + (size_t)k * cap : nullptr;
		ok = ok && e->fl[k].h_out;
	}
	ok = ok && e->dv_out;
	if (ok && hrtf) {
		const size_t T = (size_t)cfg->sh_hrir_taps;
		ok = dalloc(e->dv_hrir, kNumSh * T) && dalloc(e->dv_hist, kNumSh * T) && dalloc(e->dv_dec, 2 * (size_t)F);
		if (ok) ok = dev::h2d(e->dv_hrir, cfg->sh_hrir, sizeof(float) * kNumSh * T, e->stream) == 0;
	}
	if (ok) ok = dev::h2d(e->dv_table, encoder_table(), sizeof(float) * kTableAz * kTableEl * kShRaw, e->stream) == 0;
	if (ok) {
		// AmbisonicBinauralDecoderNode starts with num_frames_processed_on_empty_input_ = frames (no flush pending).
		DevFlags f; memset(&f, 0, sizeof(f)); f.sh_flush_frames = (int)F;
		ok = dev::h2d(e->dv_flags, &f, sizeof(f), e->stream) == 0 && dev::sync(e->stream) == 0;
	}
	if (!ok) { fprintf(stderr, "sbm_create: %s\n", dev::last_error()); sbm_destroy(e); return SBM_ERR_CUDA; }
	memset(&e->g, 0, sizeof(e->g));
	e->g.head_q[3] = 1.0f; e->g.master = 1.0f; e->g.frames = F; e->g.rate = cfg->sample_rate_hz;
	for (int k = 0; k < 2; ++k) {
		e->fl[k].done = dev::event_create();
		for (int i = 0; i < 4; ++i) e->fl[k].ev[i] = dev::event_create();
	}
	*out_engine = e;
	return SBM_OK;
}

SBM_API void sbm_destroy(sbm_engine* e) {
	if (!e) return;
	dev::use(e->cfg.device);
	if (e->stream) dev::sync(e->stream);
	SlotArrays& A = e->A;
	void* ptrs[] = {e->dv_stat, e->dv_posvol, e->dv_occl, e->dv_ring_write, e->dv_mixz, A.live, A.read_i, A.eof, A.rs, A.dgain, A.occ_cur,
	                A.lpf_prev, A.bits, A.rec, e->fl[0].dv_finished, A.mix_pos, A.mix_pos0, A.cgain, A.shg, A.shr, e->dv_flags,
	                e->dv_table, e->dv_lists, e->dv_list_ring, e->dv_bus, e->dv_partials, e->dv_out, e->dv_hrir,
	                e->dv_hist, e->dv_dec, e->dv_patches, e->dv_mix_patches};
	for (void* p : ptrs) if (p) dev::release(p);
	for (auto& b : e->buffers) if (b.owned && b.dptr) dev::release(b.dptr);
	for (auto& r : e->rings) if (r.storage) dev::release(r.storage);
	for (int k = 0; k < 2; ++k) {
		if (e->fl[k].h_out) dev::release_pinned(e->fl[k].h_out);
		if (e->fl[k].done) dev::event_destroy(e->fl[k].done);
		for (int i = 0; i < 4; ++i) if (e->fl[k].ev[i]) dev::event_destroy(e->fl[k].ev[i]);
	}
	for (void* m : e->peer_mapped) dev::peer_close(m);
	if (e->dv_gather) dev::release(e->dv_gather);
	if (e->dv_bus_sum) dev::release(e->dv_bus_sum);
	if (e->h_ring_stage) dev::release_pinned(e->h_ring_stage);
	if (e->d_ring_stage) dev::release(e->d_ring_stage);
	dev::shutdown(e->stream);
	delete e;
}

SBM_API const char* sbm_last_error(const sbm_engine* e) { return e ? e->error.c_str() : "null engine"; }

// ---------------------------------------------------------------------------------------------------------------
// PCM buffers
// ---------------------------------------------------------------------------------------------------------------
SBM_API int sbm_buffer_create(sbm_engine* e, const float* samples, size_t n, int* out_buffer) {
	if (!e || !samples || n == 0 || !out_buffer || n > 0xfffffff0ull) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	float* d = (float*)dev::alloc(sizeof(float) * (n + 8));   // slack: the mix kernel may read whole 32-byte sectors
	if (!d) return fail_cuda(e, "sbm_buffer_create");
	if (dev::h2d(d, samples, sizeof(float) * n, e->stream) || dev::sync(e->stream)) { dev::release(d); return fail_cuda(e, "sbm_buffer_create"); }
	e->buffers.push_back(BufferInfo{d, n, true});
	*out_buffer = (int)e->buffers.size() - 1;
	return SBM_OK;
}

SBM_API int sbm_buffer_adopt_device(sbm_engine* e, const float* device_base, size_t n, size_t stride_samples, int count, int* out_first_buffer) {
	if (!e || !device_base || n == 0 || count <= 0 || !out_first_buffer || n > 0xfffffff0ull) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	*out_first_buffer = (int)e->buffers.size();
	e->buffers.reserve(e->buffers.size() + (size_t)count);
	for (int k = 0; k < count; ++k) e->buffers.push_back(BufferInfo{const_cast<float*>(device_base) + (size_t)k * stride_samples, n, false});
	return SBM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// sources
// ---------------------------------------------------------------------------------------------------------------
SBM_API int sbm_source_create(sbm_engine* e, int kind, int count, int* out_first_id) {
	if (!e || count <= 0 || !out_first_id) return SBM_ERR_INVALID_ARG;
	if (kind != SBM_SOURCE_STEREO_PANNING && kind != SBM_SOURCE_BINAURAL_HQ && kind != SBM_SOURCE_NONSPATIAL) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	if (kind == SBM_SOURCE_BINAURAL_HQ && !e->dv_hrir) return fail(e, SBM_ERR_INVALID_ARG, "engine created without SH-HRIRs");
	const size_t avail = e->free_slots.size() + (size_t)(e->cap - e->n_slots);
	if ((size_t)count > avail) return fail(e, SBM_ERR_CAPACITY, "max_sources exceeded");
	*out_first_id = e->next_id;
	e->slot_of_id.reserve(e->slot_of_id.size() + (size_t)count);
	e->patches.reserve(e->patches.size() + (size_t)count);
	for (int k = 0; k < count; ++k) {
		uint32_t slot;
		if (!e->free_slots.empty()) { slot = e->free_slots.back(); e->free_slots.pop_back(); }
		else slot = e->n_slots++;
		const int id = e->next_id++;
		e->slot_of_id.push_back((int)slot);
		SrcStatic& st = e->h_stat[slot];
		st = SrcStatic{nullptr, 0, e->cfg.sample_rate_hz, kFlagAlive | kFlagLooping | ((uint32_t)kind << kFlagKindShift), 0, 0, (uint32_t)id};
		e->h_posvol[slot] = F4{0, 0, 0, 1};
		e->h_occl[slot] = 0.0f;
		e->h_ring_write[slot] = 0;
		e->d_stat.mark(slot); e->d_posvol.mark(slot); e->d_occl.mark(slot); e->d_ring.mark(slot);
		add_patch(e, SlotPatch{slot, kPatchReset | kPatchResampler, 0ull, std::numeric_limits<long long>::max(), e->cfg.sample_rate_hz, 0});
	}
	e->lists_dirty = true;
	return SBM_OK;
}

SBM_API int sbm_source_destroy(sbm_engine* e, int n, const int* ids) {
	if (!e || n < 0 || (n && !ids)) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	int rc = SBM_OK;
	for (int i = 0; i < n; ++i) {
		const int slot = slot_for(e, ids[i]);
		if (slot < 0) { rc = SBM_ERR_UNKNOWN_ID; continue; }
		kill_slot(e, (uint32_t)slot);
	}
	return rc;
}

SBM_API int sbm_source_attach_buffer(sbm_engine* e, int n, const int* ids, const int* buffers, const uint64_t* cur_read_i,
                                     const int* sampling_rate, const uint8_t* looping, const uint8_t* remove_on_finish) {
	if (!e || n < 0 || !ids || !buffers) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	int rc = SBM_OK;
	for (int i = 0; i < n; ++i) {
		const int slot = slot_for(e, ids[i]);
		if (slot < 0) { rc = SBM_ERR_UNKNOWN_ID; continue; }
		if (buffers[i] < 0 || (size_t)buffers[i] >= e->buffers.size()) { rc = SBM_ERR_INVALID_ARG; continue; }
		const int rate = sampling_rate ? sampling_rate[i] : e->cfg.sample_rate_hz;
		// AudioEngine::addSource rejects rates outside 8000..48000 (AudioEngine.cpp:884-885).
		if (rate < 8000 || rate > 48000) { rc = fail(e, SBM_ERR_INVALID_ARG, "Unsupported sampling rate for audio source"); continue; }
		const BufferInfo& b = e->buffers[(size_t)buffers[i]];
		SrcStatic& st = e->h_stat[(uint32_t)slot];
		st.pcm = b.dptr; st.len = (uint32_t)b.n; st.rate = rate;
		st.flags &= ~(kFlagLooping | kFlagRemoveOnFinish | (3u << kFlagFeedShift));
		st.flags |= (kFeedBuffer << kFlagFeedShift);
		if (!looping || looping[i]) st.flags |= kFlagLooping;
		if (remove_on_finish && remove_on_finish[i]) st.flags |= kFlagRemoveOnFinish;
		e->d_stat.mark((uint32_t)slot);
		add_patch(e, SlotPatch{(uint32_t)slot, kPatchReadI | kPatchResampler, cur_read_i ? cur_read_i[i] : 0ull, 0, rate, 0});
	}
	e->lists_dirty = true;
	return rc;
}

SBM_API int sbm_source_attach_ring(sbm_engine* e, int id, int sampling_rate, int capacity_samples) {
	if (!e || capacity_samples < 64) return SBM_ERR_INVALID_ARG;
	if (sampling_rate < 8000 || sampling_rate > 48000) return fail(e, SBM_ERR_INVALID_ARG, "Unsupported sampling rate for audio source");
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	const int slot = slot_for(e, id);
	if (slot < 0) return SBM_ERR_UNKNOWN_ID;
	RingInfo& r = e->rings[(uint32_t)slot];
	if (r.storage) { dev::sync(e->stream); dev::release(r.storage); }
	r = RingInfo{(float*)dev::alloc(sizeof(float) * ((size_t)capacity_samples + 8)), (uint32_t)capacity_samples, 0, 0};
	if (!r.storage) return fail_cuda(e, "sbm_source_attach_ring");
	SrcStatic& st = e->h_stat[(uint32_t)slot];
	st.pcm = r.storage; st.len = (uint32_t)capacity_samples; st.rate = sampling_rate;
	st.flags &= ~(3u << kFlagFeedShift);
	st.flags |= (kFeedRing << kFlagFeedShift);
	e->h_ring_write[(uint32_t)slot] = 0;
	e->d_stat.mark((uint32_t)slot); e->d_ring.mark((uint32_t)slot);
	add_patch(e, SlotPatch{(uint32_t)slot, kPatchReadI | kPatchResampler, 0ull, 0, sampling_rate, 0});
	e->lists_dirty = true;
	return SBM_OK;
}

SBM_API int sbm_source_ring_push(sbm_engine* e, int id, const float* samples, size_t n) {
	if (!e || (!samples && n)) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	const int slot = slot_for(e, id);
	if (slot < 0) return SBM_ERR_UNKNOWN_ID;
	RingInfo& r = e->rings[(uint32_t)slot];
	if (!r.storage) return fail(e, SBM_ERR_INVALID_ARG, "source is not ring-fed");
	if (r.written - r.read_known + n > r.capacity) return fail(e, SBM_ERR_CAPACITY, "ring full");
	size_t done = 0;
	while (done < n) {
		const uint32_t at = (uint32_t)((r.written + done) % r.capacity);
		const size_t chunk = std::min<size_t>(n - done, r.capacity - at);
		if (dev::h2d(r.storage + at, samples + done, sizeof(float) * chunk, e->stream)) return fail_cuda(e, "sbm_source_ring_push");
		done += chunk;
	}
	// The packet is copied from the caller's memory asynchronously only if it is pinned; make the call safe for any memory.
	if (dev::sync(e->stream)) return fail_cuda(e, "sbm_source_ring_push");
	e->pending_ring_h2d_bytes += sizeof(float) * n;
	r.written += n;
	e->h_ring_write[(uint32_t)slot] = r.written;
	e->d_ring.mark((uint32_t)slot);
	return SBM_OK;
}

// Batched form of sbm_source_ring_push: the packets of MANY ring-fed sources in one call (what the voice / streaming threads
// deliver between two buffers).  All payloads and the packet table are staged in one pinned buffer, cross PCIe as ONE
// copy, and a scatter kernel writes each packet into its source's ring -- instead of one copy + one stream
// synchronisation per source.  offsets[k] .. offsets[k + 1] is packet k inside `samples`.  Either every packet is
// accepted or none (ring full / unknown id -> nothing is written).
SBM_API int sbm_source_ring_push_batch(sbm_engine* e, int n, const int* ids, const float* samples, const size_t* offsets) {
	if (!e || n < 0 || (n && (!ids || !samples || !offsets))) return SBM_ERR_INVALID_ARG;
	if (!n) return SBM_OK;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	const size_t total = offsets[n] - offsets[0];
	std::vector<RingPush> pk((size_t)n);
	std::vector<uint32_t> slots((size_t)n);
	std::vector<unsigned long long> pending(e->rings.size(), 0);     // samples accepted earlier in this call, per slot
	for (int k = 0; k < n; ++k) {
		const int slot = slot_for(e, ids[k]);
		if (slot < 0) return SBM_ERR_UNKNOWN_ID;
		if (offsets[k + 1] < offsets[k]) return SBM_ERR_INVALID_ARG;
		RingInfo& r = e->rings[(uint32_t)slot];
		if (!r.storage) return fail(e, SBM_ERR_INVALID_ARG, "source is not ring-fed");
		const size_t m = offsets[k + 1] - offsets[k];
		unsigned long long& pend = pending[(uint32_t)slot];
		if (r.written + pend - r.read_known + m > r.capacity) return fail(e, SBM_ERR_CAPACITY, "ring full");
		pk[(size_t)k] = RingPush{r.storage, r.capacity, (uint32_t)((r.written + pend) % r.capacity), (uint32_t)(offsets[k] - offsets[0]), (uint32_t)m};
		slots[(size_t)k] = (uint32_t)slot;
		pend += m;
	}
	// staging layout: [packet table, padded to 256 B][payload]
	const size_t table_bytes = (sizeof(RingPush) * (size_t)n + 255) & ~(size_t)255;
	const size_t bytes = table_bytes + sizeof(float) * total;
	if (e->ring_stage_in_flight) { if (dev::sync(e->stream)) return fail_cuda(e, "sbm_source_ring_push_batch"); e->ring_stage_in_flight = false; }
	e->ring_stage_seq = e->buffer_index;
	if (bytes > e->ring_stage_bytes) {
		dev::sync(e->stream);
		dev::release_pinned(e->h_ring_stage); dev::release(e->d_ring_stage);
		e->ring_stage_bytes = std::max<size_t>(bytes * 2, 1 << 20);
		e->h_ring_stage = (unsigned char*)dev::alloc_pinned(e->ring_stage_bytes);
		e->d_ring_stage = (unsigned char*)dev::alloc(e->ring_stage_bytes);
		if (!e->h_ring_stage || !e->d_ring_stage) { e->ring_stage_bytes = 0; return fail_cuda(e, "sbm_source_ring_push_batch(alloc)"); }
	}
	memcpy(e->h_ring_stage, pk.data(), sizeof(RingPush) * (size_t)n);
	// A payload that already lives in page-locked memory (sbm_host_alloc_pinned, cudaHostRegister ...) is copied to the
	// device straight from the caller's buffer -- it must then stay unmodified until the next sbm_mix / sbm_mix_collect
	// returns; anything else is first copied into the engine's own pinned staging buffer.
	const bool direct = dev::is_pinned(samples + offsets[0]);
	int rc;
	if (direct) {
		rc = dev::h2d(e->d_ring_stage, e->h_ring_stage, table_bytes, e->stream) ||
		     dev::h2d(e->d_ring_stage + table_bytes, samples + offsets[0], sizeof(float) * total, e->stream);
	} else {
		memcpy(e->h_ring_stage + table_bytes, samples + offsets[0], sizeof(float) * total);
		rc = dev::h2d(e->d_ring_stage, e->h_ring_stage, bytes, e->stream);
	}
	if (rc ||
	    dev::launch_ring_scatter((const RingPush*)e->d_ring_stage, (uint32_t)n, (const float*)(e->d_ring_stage + table_bytes), e->stream))
		return fail_cuda(e, "sbm_source_ring_push_batch");
	e->ring_stage_in_flight = true;      // the pinned staging buffer is reused: the next push waits for this copy
	e->pending_ring_h2d_bytes += bytes;
	for (int k = 0; k < n; ++k) {
		RingInfo& r = e->rings[slots[(size_t)k]];
		r.written += pk[(size_t)k].n;
		e->h_ring_write[slots[(size_t)k]] = r.written;
		e->d_ring.mark(slots[(size_t)k]);
	}
	return SBM_OK;
}

// Page-locked host memory for callers that do not link CUDA themselves (packet payloads for sbm_source_ring_push_batch,
// PCM for sbm_buffer_create): copies from it need no staging.
SBM_API void* sbm_host_alloc_pinned(size_t bytes) { return dev::alloc_pinned(bytes); }
SBM_API void sbm_host_free_pinned(void* p) { dev::release_pinned(p); }

SBM_API int sbm_source_ring_size(sbm_engine* e, int id, uint64_t* out_size) {
	if (!e || !out_size) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	const int slot = slot_for(e, id);
	if (slot < 0) return SBM_ERR_UNKNOWN_ID;
	const RingInfo& r = e->rings[(uint32_t)slot];
	*out_size = r.written - r.read_known;
	return SBM_OK;
}

static int attach_mix_locked(sbm_engine* e, int id, int sampling_rate, int nz, const int* buffers, const double* pos, const double* delta, const float* mixf) {
	const int slot = slot_for(e, id);
	if (slot < 0) return SBM_ERR_UNKNOWN_ID;
	if (nz <= 0) return SBM_ERR_INVALID_ARG;
	// The mix path produces engine-rate samples; a MixSource-fed AudioSource with another rate would additionally go
	// through AudioResampler in the reference (AudioEngine.cpp:485-494).  Not needed by Substrata's callers (BikePhysics
	// uses the engine rate); rejected rather than rendered differently.
	if (sampling_rate != e->cfg.sample_rate_hz) return fail(e, SBM_ERR_UNSUPPORTED, "mix_sources at a rate other than the engine rate");
	if (e->mix_used + (uint32_t)nz > e->h_mixz.size()) return fail(e, SBM_ERR_CAPACITY, "max_mix_sources exceeded");
	for (int z = 0; z < nz; ++z) if (buffers[z] < 0 || (size_t)buffers[z] >= e->buffers.size()) return SBM_ERR_INVALID_ARG;
	const uint32_t first = e->mix_used;
	e->mix_used += (uint32_t)nz;
	for (int z = 0; z < nz; ++z) {
		const BufferInfo& b = e->buffers[(size_t)buffers[z]];
		e->h_mixz[first + (uint32_t)z] = MixZ{b.dptr, (uint32_t)b.n, mixf[z], delta[z], 0.0};
		e->d_mixz.mark(first + (uint32_t)z);
		e->mix_patches.push_back(MixPatch{first + (uint32_t)z, 0, pos[z]});
	}
	SrcStatic& st = e->h_stat[(uint32_t)slot];
	st.rate = sampling_rate; st.mix_first = first; st.nz = (uint32_t)nz;
	st.flags &= ~(3u << kFlagFeedShift);
	st.flags |= (kFeedMix << kFlagFeedShift);
	e->d_stat.mark((uint32_t)slot);
	add_patch(e, SlotPatch{(uint32_t)slot, kPatchResampler, 0ull, 0, sampling_rate, 0});
	e->lists_dirty = true;
	return SBM_OK;
}

SBM_API int sbm_source_attach_mix(sbm_engine* e, int id, int sampling_rate, int nz, const int* buffers, const double* sound_file_i,
                                  const double* source_delta, const float* mix_factor) {
	if (!e || !buffers || !sound_file_i || !source_delta || !mix_factor) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	return attach_mix_locked(e, id, sampling_rate, nz, buffers, sound_file_i, source_delta, mix_factor);
}

SBM_API int sbm_source_attach_mix1(sbm_engine* e, int n, const int* ids, const int* buffers, const double* sound_file_i,
                                   const double* source_delta, const float* mix_factor) {
	if (!e || n < 0 || !ids || !buffers || !sound_file_i || !source_delta || !mix_factor) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	e->patches.reserve(e->patches.size() + (size_t)n);
	e->mix_patches.reserve(e->mix_patches.size() + (size_t)n);
	for (int i = 0; i < n; ++i) {
		const int rc = attach_mix_locked(e, ids[i], e->cfg.sample_rate_hz, 1, buffers + i, sound_file_i + i, source_delta + i, mix_factor + i);
		if (rc) return rc;
	}
	return SBM_OK;
}

SBM_API int sbm_source_set_mix_params(sbm_engine* e, int n, const int* ids, const int* z, const double* source_delta, const float* mix_factor) {
	if (!e || n < 0 || !ids || !source_delta || !mix_factor) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	int rc = SBM_OK;
	for (int i = 0; i < n; ++i) {
		const int slot = slot_for(e, ids[i]);
		if (slot < 0) { rc = SBM_ERR_UNKNOWN_ID; continue; }
		const SrcStatic& st = e->h_stat[(uint32_t)slot];
		const uint32_t zz = z ? (uint32_t)z[i] : 0u;
		if (zz >= st.nz) { rc = SBM_ERR_INVALID_ARG; continue; }
		MixZ& m = e->h_mixz[st.mix_first + zz];
		m.delta = source_delta[i]; m.mix_factor = mix_factor[i];
		e->d_mixz.mark(st.mix_first + zz);
	}
	return rc;
}

SBM_API int sbm_source_set_eof_marker(sbm_engine* e, int id, int64_t samples_to_eof) {
	if (!e) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	const int slot = slot_for(e, id);
	if (slot < 0) return SBM_ERR_UNKNOWN_ID;
	add_patch(e, SlotPatch{(uint32_t)slot, kPatchEof, 0ull, (long long)samples_to_eof, 0, 0});
	return SBM_OK;
}

SBM_API int sbm_source_set_looping(sbm_engine* e, int id, int looping) {
	if (!e) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	const int slot = slot_for(e, id);
	if (slot < 0) return SBM_ERR_UNKNOWN_ID;
	SrcStatic& st = e->h_stat[(uint32_t)slot];
	if (looping) st.flags |= kFlagLooping; else st.flags &= ~kFlagLooping;
	e->d_stat.mark((uint32_t)slot);
	return SBM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// setters
// ---------------------------------------------------------------------------------------------------------------
SBM_API int sbm_set_source_positions(sbm_engine* e, int n, const int* ids, const float* xyz) {
	if (!e || n < 0 || (n && (!ids || !xyz))) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	int rc = SBM_OK;
	for (int i = 0; i < n; ++i) {
		const int slot = slot_for(e, ids[i]);
		if (slot < 0) { rc = SBM_ERR_UNKNOWN_ID; continue; }
		const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
		// AudioEngine drops non-finite poses before they reach Resonance (AudioEngine.cpp:951).
		if (!(std::isfinite(x) && std::isfinite(y) && std::isfinite(z))) continue;
		F4& p = e->h_posvol[(uint32_t)slot];
		p.x = x; p.y = y; p.z = z;
		e->d_posvol.mark((uint32_t)slot);
	}
	return rc;
}

SBM_API int sbm_set_source_volumes(sbm_engine* e, int n, const int* ids, const float* volume) {
	if (!e || n < 0 || (n && (!ids || !volume))) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	int rc = SBM_OK;
	for (int i = 0; i < n; ++i) {
		const int slot = slot_for(e, ids[i]);
		if (slot < 0) { rc = SBM_ERR_UNKNOWN_ID; continue; }
		e->h_posvol[(uint32_t)slot].w = volume[i];
		e->d_posvol.mark((uint32_t)slot);
	}
	return rc;
}

SBM_API int sbm_set_source_occlusions(sbm_engine* e, int n, const int* ids, const float* occ) {
	if (!e || n < 0 || (n && (!ids || !occ))) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	int rc = SBM_OK;
	for (int i = 0; i < n; ++i) {
		const int slot = slot_for(e, ids[i]);
		if (slot < 0) { rc = SBM_ERR_UNKNOWN_ID; continue; }
		e->h_occl[(uint32_t)slot] = occ[i];
		e->d_occl.mark((uint32_t)slot);
	}
	return rc;
}

SBM_API int sbm_set_head(sbm_engine* e, const float* pos, const float* q) {
	if (!e || !pos || !q) return SBM_ERR_INVALID_ARG;
	for (int i = 0; i < 3; ++i) if (!std::isfinite(pos[i])) return SBM_OK;   // AudioEngine.cpp:985
	for (int i = 0; i < 4; ++i) if (!std::isfinite(q[i])) return SBM_OK;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	memcpy(e->g.head_pos, pos, 12);
	memcpy(e->g.head_q, q, 16);
	return SBM_OK;
}

SBM_API int sbm_set_master_volume(sbm_engine* e, float volume) {
	if (!e) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	e->g.master = volume;
	return SBM_OK;
}

SBM_API int sbm_enable_room_effects(sbm_engine* e, int enable) {
	if (!e) return SBM_ERR_INVALID_ARG;
	return enable ? fail(e, SBM_ERR_UNSUPPORTED, "room effects are not on the accelerated path (SURVEY.md 8f)") : SBM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// the per-buffer call
// ---------------------------------------------------------------------------------------------------------------
// One buffer = sbm_mix_begin (uploads, K1, K2, K3) -> [the caller may sum buses across ranks here] -> the finish part
// (gather, K4, D2H of the result into pinned memory, an event) -> collection (wait for the event, hand the samples out,
// host bookkeeping).  sbm_mix / sbm_mix_finish do all of it synchronously; sbm_mix_submit + sbm_mix_collect split it so
// that up to two buffers are in flight and the host's launch latency hides behind the previous buffer's kernels -- the
// mixer thread of the reference also runs ahead of the device callback (AudioEngine.cpp:540-545).
static int mix_begin_locked(sbm_engine* e) {
	if (e->begun) return fail(e, SBM_ERR_INVALID_ARG, "sbm_mix_begin called twice without sbm_mix_finish");
	sbm_engine::InFlight& fl = e->fl[e->buffer_index & 1];
	if (fl.pending) return fail(e, SBM_ERR_INVALID_ARG, "two buffers already in flight: call sbm_mix_collect first");
	// bytes pushed into rings since the previous buffer belong to this buffer's host->device traffic
	e->stats.last_h2d_bytes = e->pending_ring_h2d_bytes; e->stats.last_d2h_bytes = 0;
	e->pending_ring_h2d_bytes = 0;
	void* st = e->stream;
	// 1. pending parameter uploads
	if (upload_range(e, e->dv_stat, e->h_stat, e->d_stat) || upload_range(e, e->dv_posvol, e->h_posvol, e->d_posvol) ||
	    upload_range(e, e->dv_occl, e->h_occl, e->d_occl) || upload_range(e, e->dv_ring_write, e->h_ring_write, e->d_ring) ||
	    upload_range(e, e->dv_mixz, e->h_mixz, e->d_mixz))
		return fail_cuda(e, "sbm_mix_begin(upload)");
	if (e->lists_dirty) {
		rebuild_lists(e);
		if ((!e->list_pan.empty() && dev::h2d(e->dv_lists, e->list_pan.data(), 4 * e->list_pan.size(), st)) ||
		    (!e->list_sh.empty() && dev::h2d(e->dv_lists + e->list_pan.size(), e->list_sh.data(), 4 * e->list_sh.size(), st)) ||
		    (!e->list_ring.empty() && dev::h2d(e->dv_list_ring, e->list_ring.data(), 4 * e->list_ring.size(), st)))
			return fail_cuda(e, "sbm_mix_begin(lists)");
		e->stats.last_h2d_bytes += 4 * (e->list_pan.size() + e->list_sh.size() + e->list_ring.size());
	}
	if (!e->patches.empty() || !e->mix_patches.empty()) {
		if (upload_list(e, e->dv_patches, e->dv_patches_cap, e->patches) || upload_list(e, e->dv_mix_patches, e->dv_mix_patches_cap, e->mix_patches))
			return fail_cuda(e, "sbm_mix_begin(patches)");
		if (dev::launch_patch(e->A, e->dv_patches, (uint32_t)e->patches.size(), e->dv_mix_patches, (uint32_t)e->mix_patches.size(), st))
			return fail_cuda(e, "k_patch");
		// the host vectors must stay valid until the copy has been consumed (pageable source => staged synchronously)
		for (const SlotPatch& p : e->patches) e->patch_of_slot[p.slot] = -1;
		e->patches.clear(); e->mix_patches.clear();
	}
	e->g.parity = (uint32_t)(e->buffer_index & 1);
	e->g.n_slots = e->n_slots;
	e->g.n_pan = (uint32_t)e->list_pan.size(); e->g.n_sh = (uint32_t)e->list_sh.size();
	e->A.finished = fl.dv_finished;      // per-parity list: the previous buffer's may not have been collected yet
	// 2. kernels
	if (e->profiling) dev::event_record(fl.ev[0], st);
	if (e->g.n_pan + e->g.n_sh && dev::launch_params(e->g, e->A, e->dv_flags, e->dv_table, e->dv_lists, e->n_slots, st)) return fail_cuda(e, "k1_params");
	if (e->profiling) dev::event_record(fl.ev[1], st);
	MixArgs m;
	m.rec = e->A.rec; m.shg = e->A.shg; m.shr = e->A.shr; m.mixz = e->dv_mixz; m.mix_pos0 = e->A.mix_pos0; m.rs = e->A.rs; m.lpf_prev = e->A.lpf_prev;
	m.n_pan = (uint32_t)e->list_pan.size();
	m.n_sh = (uint32_t)e->list_sh.size();
	m.frames = e->g.frames; m.rate = e->g.rate; m.partials = e->dv_partials; m.partials_floats = e->partials_floats;
	m.bus = e->dv_bus; m.flags = e->dv_flags; m.parity = e->g.parity;
	m.ev_after_k2 = e->profiling ? fl.ev[2] : nullptr;
	m.peer = e->peer; m.bus_floats = e->bus_floats;
	if (dev::launch_mix(m, st)) return fail_cuda(e, "k2_mix");
	e->begun = true;
	fl.active = (uint32_t)(e->list_pan.size() + e->list_sh.size());
	fl.h2d_bytes = e->stats.last_h2d_bytes;
	return SBM_OK;
}

// Enqueues the finish part of the buffer begun last; no host synchronisation.
static int mix_enqueue_finish_locked(sbm_engine* e) {
	if (!e->begun) return fail(e, SBM_ERR_INVALID_ARG, "sbm_mix_finish without sbm_mix_begin");
	e->begun = false;
	void* st = e->stream;
	const uint32_t F = e->g.frames;
	const uint32_t parity = (uint32_t)(e->buffer_index & 1);
	sbm_engine::InFlight& fl = e->fl[parity];
	if (e->peer.world) {
		// wait for every rank's bus of this buffer (K3 stored them into our gather area) and sum them in rank order
		const unsigned int target = (unsigned int)(e->peer.world * ((e->buffer_index >> 1) - (e->peer_epoch >> 1) + 1));
		if (dev::launch_peer_gather(e->peer, e->dv_bus_sum, e->bus_floats, parity, target, st)) return fail_cuda(e, "k_peer_gather");
	}
	FinishArgs f;
	f.bus = e->peer.world ? e->dv_bus_sum : e->dv_bus; f.hrir = e->dv_hrir; f.taps = (uint32_t)e->cfg.sh_hrir_taps; f.hist = e->dv_hist; f.dec = e->dv_dec;
	f.flags = e->dv_flags; f.parity = parity; f.frames = F; f.out = e->dv_out;
	f.list_ring = e->dv_list_ring; f.n_ring = (uint32_t)e->list_ring.size(); f.read_i = e->A.read_i; f.finished_count_reset = nullptr; f.hist_smem_all = 0;
	if (dev::launch_finish(f, st)) return fail_cuda(e, "k4_finish");
	if (e->profiling) dev::event_record(fl.ev[3], st);
	const size_t bytes = sizeof(float) * 2 * F + sizeof(OutTail) + sizeof(unsigned long long) * e->list_ring.size();
	if (dev::d2h(fl.h_out, e->dv_out, bytes, st)) return fail_cuda(e, "sbm_mix_finish(d2h)");
	dev::event_record(fl.done, st);
	fl.pending = true; fl.seq = e->buffer_index; fl.list_ring = e->list_ring; fl.d2h_bytes = bytes;
	e->buffer_index++;
	return SBM_OK;
}

// Waits for the oldest buffer in flight and hands it out.
static int mix_collect_locked(sbm_engine* e, float* out_interleaved, int* out_valid) {
	sbm_engine::InFlight& fl = e->fl[e->collected & 1];
	if (!fl.pending || fl.seq != e->collected) return fail(e, SBM_ERR_INVALID_ARG, "sbm_mix_collect: no buffer in flight");
	void* st = e->stream;
	const uint32_t F = e->g.frames;
	if (dev::event_sync(fl.done)) return fail_cuda(e, "sbm_mix_collect");
	fl.pending = false;
	e->collected++;
	if (e->ring_stage_in_flight && e->ring_stage_seq <= fl.seq) e->ring_stage_in_flight = false;
	e->stats.last_h2d_bytes = fl.h2d_bytes;
	e->stats.last_d2h_bytes = fl.d2h_bytes;
	e->stats.active_sources = fl.active;
	OutTail tail;
	memcpy(&tail, fl.h_out + sizeof(float) * 2 * F, sizeof(tail));
	if (out_interleaved) memcpy(out_interleaved, fl.h_out, sizeof(float) * 2 * F);
	if (out_valid) *out_valid = tail.valid;
	e->last_valid = tail.valid;
	// ring read counters
	const unsigned long long* reads = (const unsigned long long*)(fl.h_out + sizeof(float) * 2 * F + sizeof(OutTail));
	for (size_t k = 0; k < fl.list_ring.size(); ++k) {
		RingInfo& r = e->rings[fl.list_ring[k]];
		if (r.storage && reads[k] <= r.written && reads[k] >= r.read_known) r.read_known = reads[k];     // (the slot may have been recycled since)
	}
	// sources stopped by the feeder (AudioEngine.cpp:499-515)
	if (tail.n_finished > 0) {
		std::vector<int> ids((size_t)tail.n_finished);
		if (dev::d2h(ids.data(), fl.dv_finished, sizeof(int) * ids.size(), st) || dev::sync(st)) return fail_cuda(e, "sbm_mix_finish(finished)");
		e->stats.last_d2h_bytes += sizeof(int) * ids.size();
		std::sort(ids.begin(), ids.end());
		for (int id : ids) {
			const int slot = slot_for(e, id);
			if (slot >= 0) kill_slot(e, (uint32_t)slot);
			e->finished_queue.push_back(id);
		}
	}
	e->stats.buffers_mixed++;
	e->stats.kernel_launches = dev::launches();
	e->stats.last_algorithmic_bytes = tail.algo_bytes;
	e->stats.last_rendered_sources = tail.n_rendered;
	if (e->profiling) {
		e->stats.last_params_kernel_ms = dev::event_elapsed_ms(fl.ev[0], fl.ev[1]);
		e->stats.last_mix_kernel_ms = dev::event_elapsed_ms(fl.ev[1], fl.ev[2]);
		e->stats.last_finish_kernel_ms = dev::event_elapsed_ms(fl.ev[2], fl.ev[3]);
	}
	return SBM_OK;
}

SBM_API int sbm_mix_begin(sbm_engine* e) {
	if (!e) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	if (e->collected != e->buffer_index) return fail(e, SBM_ERR_INVALID_ARG, "sbm_mix_begin with a submitted buffer outstanding: use sbm_mix_submit / sbm_mix_collect");
	return mix_begin_locked(e);
}

SBM_API int sbm_mix_finish(sbm_engine* e, float* out_interleaved, int* out_valid) {
	if (!e) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	const int rc = mix_enqueue_finish_locked(e);
	if (rc) return rc;
	return mix_collect_locked(e, out_interleaved, out_valid);
}

SBM_API int sbm_mix(sbm_engine* e, float* out_interleaved, int* out_valid) {
	if (!e || !out_interleaved) return SBM_ERR_INVALID_ARG;
	const int rc = sbm_mix_begin(e);
	if (rc) return rc;
	return sbm_mix_finish(e, out_interleaved, out_valid);
}

SBM_API int sbm_mix_submit(sbm_engine* e) {
	if (!e) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	const int rc = mix_begin_locked(e);
	if (rc) return rc;
	return mix_enqueue_finish_locked(e);
}

SBM_API int sbm_mix_collect(sbm_engine* e, float* out_interleaved, int* out_valid) {
	if (!e) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	return mix_collect_locked(e, out_interleaved, out_valid);
}

// ---------------------------------------------------------------------------------------------------------------
// multi-GPU: one engine per GPU / process, sources sharded, buses exchanged over NVLink peer memory
// ---------------------------------------------------------------------------------------------------------------
static size_t gather_area_floats(const sbm_engine* e) { return 2 * (size_t)kMaxPeers * e->bus_floats; }

SBM_API int sbm_peer_export(sbm_engine* e, void* handle_out, size_t handle_bytes) {
	if (!e || !handle_out || handle_bytes < 64) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	if (!e->dv_gather) {
		// area + 16 counters' worth of padding: [2 arrival counters][done counter]
		e->dv_gather = (float*)dev::alloc(sizeof(float) * (gather_area_floats(e) + 16));
		e->dv_bus_sum = (float*)dev::alloc(sizeof(float) * e->bus_floats);
		if (!e->dv_gather || !e->dv_bus_sum) return fail_cuda(e, "sbm_peer_export(alloc)");
	}
	if (dev::peer_export(e->dv_gather, handle_out)) return fail_cuda(e, "sbm_peer_export");
	return SBM_OK;
}

SBM_API int sbm_peer_connect(sbm_engine* e, int rank, int world, const void* handles, size_t handle_stride) {
	if (!e || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || (world > 1 && (!handles || handle_stride < 64))) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	if (!e->dv_gather) return fail(e, SBM_ERR_INVALID_ARG, "sbm_peer_connect before sbm_peer_export");
	if (e->begun || e->collected != e->buffer_index) return fail(e, SBM_ERR_INVALID_ARG, "sbm_peer_connect with a buffer in flight");
	if (e->peer.world) return fail(e, SBM_ERR_INVALID_ARG, "already connected");
	// the arrival counters count buffers per parity from now on: all ranks must connect at the same, even, buffer index
	if (e->buffer_index & 1) return fail(e, SBM_ERR_INVALID_ARG, "sbm_peer_connect needs an even number of buffers mixed so far (all ranks alike)");
	PeerArgs pa{};
	pa.rank = (uint32_t)rank; pa.world = (uint32_t)world;
	for (int p = 0; p < world; ++p) {
		float* base = e->dv_gather;
		if (p != rank) {
			e->peer_mapped[p] = dev::peer_open((const unsigned char*)handles + (size_t)p * handle_stride);
			if (!e->peer_mapped[p]) return fail_cuda(e, "sbm_peer_connect");
			base = (float*)e->peer_mapped[p];
		}
		pa.gather[p] = base;
		pa.arrive[p] = (unsigned int*)(base + gather_area_floats(e));
	}
	pa.done = (unsigned int*)(e->dv_gather + gather_area_floats(e)) + 4;
	e->peer = pa;
	e->peer_epoch = e->buffer_index;
	return SBM_OK;
}

SBM_API int sbm_bus_device(sbm_engine* e, float** out_device_ptr, size_t* out_num_floats) {
	if (!e || !out_device_ptr || !out_num_floats) return SBM_ERR_INVALID_ARG;
	*out_device_ptr = e->dv_bus; *out_num_floats = e->bus_floats;
	return SBM_OK;
}
SBM_API int sbm_out_device(sbm_engine* e, float** out_device_ptr) {
	if (!e || !out_device_ptr) return SBM_ERR_INVALID_ARG;
	*out_device_ptr = e->dv_out;
	return SBM_OK;
}
SBM_API void* sbm_stream(sbm_engine* e) { return e ? e->stream : nullptr; }

SBM_API int sbm_poll_finished(sbm_engine* e, int* ids, int cap, int* out_n) {
	if (!e || !ids || cap < 0 || !out_n) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	const int n = std::min<int>(cap, (int)e->finished_queue.size());
	for (int i = 0; i < n; ++i) ids[i] = e->finished_queue[(size_t)i];
	e->finished_queue.erase(e->finished_queue.begin(), e->finished_queue.begin() + n);
	*out_n = n;
	return SBM_OK;
}

SBM_API int sbm_source_query(sbm_engine* e, int id, sbm_source_state* out) {
	if (!e || !out) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	memset(out, 0, sizeof(*out));
	const int slot = slot_for(e, id);
	if (slot < 0) return SBM_ERR_UNKNOWN_ID;
	const SrcStatic& st = e->h_stat[(uint32_t)slot];
	unsigned long long rd = 0; long long eof = 0; float dg = 0, occ = 0; double pos = 0; uint32_t live = 0;
	int rc = dev::d2h(&rd, e->A.read_i + slot, 8, e->stream) | dev::d2h(&eof, e->A.eof + slot, 8, e->stream) |
	         dev::d2h(&dg, e->A.dgain + slot, 4, e->stream) | dev::d2h(&occ, e->A.occ_cur + slot, 4, e->stream) |
	         dev::d2h(&live, e->A.live + slot, 4, e->stream);
	if (st.nz) rc |= dev::d2h(&pos, e->A.mix_pos + st.mix_first, 8, e->stream);
	if (rc || dev::sync(e->stream)) return fail_cuda(e, "sbm_source_query");
	const bool ring = ((st.flags >> kFlagFeedShift) & 3) == kFeedRing;
	out->cur_read_i = ring ? 0 : rd;
	out->ring_size = ring ? e->rings[(uint32_t)slot].written - rd : 0;
	out->eof_marker = eof; out->mix_pos0 = pos; out->direct_gain = dg; out->occlusion = occ; out->alive = (int)live;
	return SBM_OK;
}

SBM_API int sbm_get_stats(sbm_engine* e, sbm_stats* out) {
	if (!e || !out) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	e->stats.kernel_launches = dev::launches();
	*out = e->stats;
	return SBM_OK;
}

SBM_API int sbm_set_profiling(sbm_engine* e, int enable) {
	if (!e) return SBM_ERR_INVALID_ARG;
	e->profiling = enable != 0;
	return SBM_OK;
}

SBM_API float sbm_debug_ramp_advance(float g, float inc, uint32_t n) { return ramp_advance(g, inc, n); }
SBM_API float sbm_debug_atan2f(float y, float x) { return exact_atan2f(y, x); }
SBM_API const float* sbm_debug_encoder_table(void) { return encoder_table(); }
SBM_API int sbm_debug_get_ambisonic_bus(sbm_engine* e, float* out16xF) {
	if (!e || !out16xF) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	if (dev::d2h(out16xF, e->dv_bus, sizeof(float) * kNumSh * e->g.frames, e->stream) || dev::sync(e->stream)) return fail_cuda(e, "debug bus");
	return SBM_OK;
}

}  // extern "C"
