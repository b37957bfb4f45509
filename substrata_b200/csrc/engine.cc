// engine.cc -- host side of the B200 mixer behind the C ABI of include/sbm_audio.h.
//
// Owns: the id -> slot table, host mirrors of the host-written per-slot parameters (with dirty ranges so that one
// sbm_mix uploads only what the setters touched), the PCM buffer registry, ring storage for streamed sources, and
// the per-buffer launch sequence  patch -> K1 (params) -> K2 (mix) -> K3 (reduce) -> [all-reduce by the caller] ->
// K4 (decode + stereo mix + interleave) -> D2H.  All arithmetic on audio happens in the kernels (device_api.h);
// nothing here touches a sample except to copy pushed ring packets to the device.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sbm_audio.h"
#include "device_api.h"

// NVTX ranges around the host phases of a buffer (visible in Nsight Systems next to the kernels; no-ops without a tool attached).
#if defined(__has_include)
#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
#define SBM_HAVE_NVTX 1
#endif
#endif
namespace {
struct NvtxRange {
#ifdef SBM_HAVE_NVTX
	explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
	~NvtxRange() { nvtxRangePop(); }
#else
	explicit NvtxRange(const char*) {}
#endif
	NvtxRange(const NvtxRange&) = delete;
	NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace

namespace sbm {
const float* encoder_table();   // sh_table.cc
}

using namespace sbm;

namespace {

// Which elements of a host-written per-slot array changed since the last buffer: an index list (deduplicated by an epoch
// stamp per element) for the usual few-thousand scattered setters per frame, or a dense range once a batched setter has
// touched a whole run of slots.
struct DirtySet {
	std::vector<uint32_t> idx, stamp;
	uint32_t epoch = 1, lo = 0xffffffffu, hi = 0;
	bool dense = false;
	void init(size_t cap) { stamp.assign(cap, 0u); }
	void mark(uint32_t i) {
		if (i < lo) lo = i;
		if (i + 1 > hi) hi = i + 1;
		if (!dense && stamp[i] != epoch) { stamp[i] = epoch; idx.push_back(i); }
	}
	void mark_range(uint32_t a, uint32_t b) { if (b <= a) return; if (a < lo) lo = a; if (b > hi) hi = b; dense = true; }
	bool any() const { return hi > lo; }
	void clear() {
		idx.clear(); lo = 0xffffffffu; hi = 0; dense = false;
		if (++epoch == 0) { std::fill(stamp.begin(), stamp.end(), 0u); epoch = 1; }
	}
};

struct BufferInfo { float* dptr; size_t n; bool owned; uint32_t refs; bool doomed; };

struct RingInfo { float* storage; uint32_t capacity; unsigned long long written; unsigned long long read_known; };

// One setter call, queued by any thread without taking a lock (Resonance posts its setters to a lock-free task queue the
// same way: RA/utils/lockless_task_queue.cc:43-51, RA/graph/resonance_audio_api_impl.cc:305-317) and applied by the mixer
// thread at the next sbm_mix_begin / sbm_mix_submit.  Payload follows the header.
struct SetterRec {
	SetterRec* next;
	uint32_t kind, n;
};
enum : uint32_t { kSetPositions, kSetVolumes, kSetOcclusions, kSetHead, kSetMaster, kSetMixParams };

// Page-locked staging for everything the host sends per buffer (parameter upload blocks, ring packets), with a device
// mirror of the same size: a FIFO ring -- regions are allocated in stream order and released when the buffer that follows
// them has been collected.  Every H2D the engine issues per buffer reads from here (or from memory the caller pinned), so
// no copy ever synchronises the stream behind the caller's back.
struct StageRing {
	unsigned char *h = nullptr, *d = nullptr;
	size_t cap = 0;
	unsigned long long head = 0, tail = 0;      // monotonic byte counters
	uint32_t gen = 0;
	size_t alloc(size_t bytes) {                  // offset of a 256-byte aligned region, or (size_t)-1
		bytes = (bytes + 255) & ~(size_t)255;
		if (!cap || bytes > cap) return (size_t)-1;
		unsigned long long at = head;
		const size_t pos = (size_t)(at % cap);
		if (pos + bytes > cap) at += cap - pos;
		if (at + bytes - tail > cap) return (size_t)-1;
		head = at + bytes;
		return (size_t)(at % cap);
	}
};
struct Retired { void* pinned; void* device; unsigned long long after_seq; };   // freed once buffer `after_seq` has been collected

}  // namespace

struct sbm_engine {
	sbm_config cfg;
	void* stream = nullptr;
	void* upload_stream = nullptr;   // sbm_buffer_create: PCM uploads do not queue behind (or hold up) the mixer's stream
	void* copy_stream = nullptr;     // the result's D2H: behind an event after K4b, so that the next buffer's K1 does not queue behind the copy
	void* early_stream = nullptr;    // the next buffer's position / occlusion upload, while the previous buffer's mix kernel runs (upload_block_locked)
	void* ev_k1_done = nullptr;      // after K1 on the mixer's stream: K1 is the only reader of what an early upload writes
	void* ev_early = nullptr;        // after an early upload: what this buffer's K1 waits for
	bool k1_event_valid = false;
	std::string error;
	std::mutex mu;               // guards the engine state below; setters never take it, and it is never held across a wait on the GPU
	std::atomic<SetterRec*> setq{nullptr};
	std::atomic<unsigned long long> setters_queued{0};

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
	DirtySet d_stat, d_posvol, d_occl, d_ring, d_mixz;
	uint32_t mix_used = 0;       // MixZ entries handed out so far (high-water mark; freed ranges are recycled through mix_free)
	std::map<uint32_t, std::vector<uint32_t>> mix_free;   // nz -> first entries of free MixZ ranges of that length
	std::vector<int> slot_buf;   // slot -> PCM buffer its feeder references (-1 = none)
	std::vector<int> mixz_buf;   // MixZ entry -> PCM buffer
	std::vector<SlotPatch> patches;          // at most one (merged) pending patch per slot
	std::vector<int> patch_of_slot;          // slot -> index into patches, -1 = none pending
	std::vector<MixPatch> mix_patches;
	std::vector<RingInfo> rings;            // indexed by slot (storage == nullptr when not a ring source)
	std::vector<uint32_t> list_pan, list_sh, list_ring, list_mix_pan, list_mix_sh;     // list_mix_*: SrcStatic::mix_first per render-list entry (K1 hint)
	bool lists_dirty = true;
	std::vector<BufferInfo> buffers;
	std::vector<int> finished_queue;
	std::vector<Retired> retired;

	// device
	SlotArrays A;                // device pointers
	SrcStatic* dv_stat = nullptr; F4* dv_posvol = nullptr; float* dv_occl = nullptr; unsigned long long* dv_ring_write = nullptr;
	MixZ* dv_mixz = nullptr;
	DevFlags* dv_flags = nullptr;
	float* dv_table = nullptr;
	uint32_t *dv_lists = nullptr, *dv_list_ring = nullptr, *dv_list_mix = nullptr;   // dv_lists: pan list followed by the ambisonic list
	float* dv_partials = nullptr; size_t partials_floats = 0;
	float* dv_bus = nullptr; size_t bus_floats = 0;
	float* dv_hrir = nullptr; float* dv_hist = nullptr; float* dv_dec = nullptr;
	float* dv_out = nullptr; size_t out_bytes = 0;     // TWO result blocks (one per buffer in flight), out_bytes each
	StageRing stage;
	// multi-GPU bus exchange over peer memory (sbm_peer_export / sbm_peer_connect)
	PeerArgs peer{};                         // world == 0 until connected
	float* dv_gather = nullptr;              // own gather area [2][kMaxPeers][bus_floats] followed by the counters
	float* dv_bus_sum = nullptr;             // bus summed over ranks (what K4 reads when connected)
	void* peer_mapped[kMaxPeers] = {};       // IPC mappings of the other ranks' areas
	unsigned long long peer_buffers = 0;     // collective buffers since sbm_peer_connect / sbm_peer_resync
	unsigned long long peer_timeout_ns = 2000000000ull;
	size_t pending_ring_h2d_bytes = 0;       // ring pushes since the last sbm_mix_begin (reported with the next buffer's stats)
	// Up to two buffers in flight (sbm_mix_submit / sbm_mix_collect): everything a submitted buffer's completion needs.
	struct InFlight {
		bool pending = false;
		unsigned long long seq = 0;          // buffer_index of the buffer
		unsigned char* h_out = nullptr;      // pinned D2H staging: 2F floats + OutTail + ring read counters + inline finished ids
		void* done = nullptr;                // event after the D2H (on copy_stream)
		void* k4_done = nullptr;             // event after K4b (on the mixer's stream): what the D2H waits for
		void* ev[4] = {nullptr, nullptr, nullptr, nullptr};   // profiling: before K1 / after K1 / after K2 / after K4
		int* dv_finished = nullptr;          // this parity's finished-id list on the device
		std::vector<uint32_t> list_ring;     // ring-fed slots in the order their read counters were reported
		size_t h2d_bytes = 0, d2h_bytes = 0;
		uint32_t active = 0;
		unsigned long long stage_mark = 0; uint32_t stage_gen = 0;   // staging consumed up to here once this buffer is done
	} fl[2];
	unsigned long long collected = 0;        // buffers collected so far (== buffer_index when nothing is in flight)
	Globals g;
	unsigned long long buffer_index = 0;
	bool begun = false;
	int last_valid = 0;
	// room effects (RA/graph/graph_manager.cc:106-107,144-151,182-198).  System settings as the setters left
	// them, the ReflectionsNode's cached copy (what its Update compares against), and the device side of the node.
	struct Room {
		bool configured = false;             // sbm_set_room_sh_hrir was called: the room kernel and the order-1 decoder exist
		bool enabled = false;
		float pos[3] = {0, 0, 0}, rot[4] = {0, 0, 0, 1}, dims[3] = {0, 0, 0}, coeffs[6] = {0, 0, 0, 0, 0, 0}, cutoff = 0, gain = 0;   // ReflectionProperties()
		float n_pos[3] = {0, 0, 0}, n_rot[4] = {0, 0, 0, 1}, n_dims[3] = {0, 0, 0}, n_coeffs[6] = {0, 0, 0, 0, 0, 0}, n_cutoff = 0, n_listener[3] = {0, 0, 0};
		uint32_t tail_buffers = 0;           // buffers the node may still need to play out its delay line after the sends were muted
		uint32_t taps1 = 0;
		uint32_t delay_size = 0;
		RoomParams* dv_params = nullptr;     // [2]: one per buffer parity
		RoomState* dv_state = nullptr;
		float* dv_delay = nullptr; float* dv_foa = nullptr; float* dv_hrir1 = nullptr; float* dv_hist1 = nullptr; float* dv_dec1 = nullptr;
		RoomParams h_params;                 // this buffer's parameters (staged through the upload block)
		bool run_this_buffer = false;
	} room;

	// stats
	bool profiling = false;
	int overlap_mode = 1;
	bool overlap_allowed = false;   // K1 beside K2 for large stereo-bus scenes (mix_begin_locked): SBM_OVERLAP=1 in the environment turns it on
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

constexpr uint32_t kOverlapMinSources = 32768;     // below this K1 is a single wave: nothing to hide

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
	e->list_pan.clear(); e->list_sh.clear(); e->list_ring.clear(); e->list_mix_pan.clear(); e->list_mix_sh.clear();
	for (uint32_t s = 0; s < e->n_slots; ++s) {
		const SrcStatic& st = e->h_stat[s];
		if (!(st.flags & kFlagAlive)) continue;
		const uint32_t kind = (st.flags >> kFlagKindShift) & 0xff;
		const uint32_t mix = st.nz ? st.mix_first : 0xffffffffu;      // K1 requests the first MixZ together with the slot's state
		if (kind == kKindHrtf) { e->list_sh.push_back(s); e->list_mix_sh.push_back(mix); } else { e->list_pan.push_back(s); e->list_mix_pan.push_back(mix); }
		if (((st.flags >> kFlagFeedShift) & 3) == kFeedRing) e->list_ring.push_back(s);
	}
	// entry k of [pan list | ambisonic list] == slot k + base (one run of sources, nothing removed in between): K1 needs no list load
	const size_t np = e->list_pan.size(), ns = e->list_sh.size();
	bool ident = np + ns > 0;
	const uint32_t base = np ? e->list_pan[0] : (ns ? e->list_sh[0] : 0u);
	for (size_t k = 0; k < np && ident; ++k) ident = e->list_pan[k] == base + (uint32_t)k;
	for (size_t k = 0; k < ns && ident; ++k) ident = e->list_sh[k] == base + (uint32_t)(np + k);
	e->g.list_identity = ident ? 1u : 0u; e->g.list_base = base;
	bool mident = np + ns > 0;
	const uint32_t mbase = np ? e->list_mix_pan[0] : (ns ? e->list_mix_sh[0] : 0u);
	if (mbase == 0xffffffffu) mident = false;
	for (size_t k = 0; k < np && mident; ++k) mident = e->list_mix_pan[k] == mbase + (uint32_t)k;
	for (size_t k = 0; k < ns && mident; ++k) mident = e->list_mix_sh[k] == mbase + (uint32_t)(np + k);
	e->g.mix_identity = mident ? 1u : 0u; e->g.mix_base = mbase;
	e->lists_dirty = false;
}

// Device (and pinned) memory that kernels already enqueued may still touch is not freed on the spot: it is parked until
// the next buffer to be submitted has been collected (stream order: everything enqueued before it is done by then).
void retire(sbm_engine* e, void* pinned, void* device) {
	if (pinned || device) e->retired.push_back(Retired{pinned, device, e->buffer_index});
}
void free_retired(sbm_engine* e, bool all) {
	size_t k = 0;
	for (size_t i = 0; i < e->retired.size(); ++i) {
		const Retired& r = e->retired[i];
		if (all || e->collected > r.after_seq) { if (r.pinned) dev::release_pinned(r.pinned); if (r.device) dev::release(r.device); }
		else e->retired[k++] = r;
	}
	e->retired.resize(k);
}

void buffer_unref(sbm_engine* e, int buf) {
	if (buf < 0 || (size_t)buf >= e->buffers.size()) return;
	BufferInfo& b = e->buffers[(size_t)buf];
	if (b.refs) --b.refs;
	if (!b.refs && b.doomed && b.dptr) { if (b.owned) retire(e, nullptr, b.dptr); b.dptr = nullptr; b.n = 0; }
}

// Drops what a slot's feeder references: its PCM buffer, its MixZ range (back to the free list of that length) and the buffers
// of its mix sources.  The reference frees an AudioSource's mix_sources and shared_buffer reference with the source.
void detach_slot_refs(sbm_engine* e, uint32_t slot) {
	if (e->slot_buf[slot] >= 0) { buffer_unref(e, e->slot_buf[slot]); e->slot_buf[slot] = -1; }
	SrcStatic& st = e->h_stat[slot];
	if (st.nz) {
		for (uint32_t z = 0; z < st.nz; ++z) { buffer_unref(e, e->mixz_buf[st.mix_first + z]); e->mixz_buf[st.mix_first + z] = -1; }
		e->mix_free[st.nz].push_back(st.mix_first);
		st.nz = 0; st.mix_first = 0;
	}
}

void kill_slot(sbm_engine* e, uint32_t slot) {
	SrcStatic& st = e->h_stat[slot];
	if (!(st.flags & kFlagAlive)) return;
	if (st.id < e->slot_of_id.size()) e->slot_of_id[st.id] = -1;
	st.flags &= ~kFlagAlive;
	detach_slot_refs(e, slot);
	e->d_stat.mark(slot);
	if (e->rings[slot].storage) { retire(e, nullptr, e->rings[slot].storage); e->rings[slot] = RingInfo{nullptr, 0, 0, 0}; }
	e->free_slots.push_back(slot);
	e->lists_dirty = true;
}

// ---- staging ----
// A region of `bytes` in the staging ring; grows the ring (the old one is retired, never waited for) when it is full.
int stage_alloc(sbm_engine* e, size_t bytes, size_t* off) {
	size_t at = e->stage.alloc(bytes);
	if (at == (size_t)-1) {
		const size_t want = std::max<size_t>(std::max<size_t>(e->stage.cap * 2, 4 * bytes + 4096), (size_t)1 << 20);
		unsigned char* h = (unsigned char*)dev::alloc_pinned(want);
		unsigned char* d = (unsigned char*)dev::alloc(want);
		if (!h || !d) { dev::release_pinned(h); dev::release(d); return -1; }
		retire(e, e->stage.h, e->stage.d);
		e->stage.h = h; e->stage.d = d; e->stage.cap = want; e->stage.head = e->stage.tail = 0; e->stage.gen++;
		at = e->stage.alloc(bytes);
		if (at == (size_t)-1) return -1;
	}
	*off = at;
	return 0;
}

// ---- setter queue ----
SetterRec* setter_new(uint32_t kind, uint32_t n, size_t payload_bytes) {
	SetterRec* r = (SetterRec*)malloc(sizeof(SetterRec) + payload_bytes);
	if (r) { r->next = nullptr; r->kind = kind; r->n = n; }
	return r;
}
inline unsigned char* setter_payload(SetterRec* r) { return (unsigned char*)(r + 1); }
void setter_push(sbm_engine* e, SetterRec* r) {
	r->next = e->setq.load(std::memory_order_relaxed);
	while (!e->setq.compare_exchange_weak(r->next, r, std::memory_order_release, std::memory_order_relaxed)) {}
	e->setters_queued.fetch_add(1, std::memory_order_relaxed);
}

// ids[0..n) ascending by one AND their slots ascending by one (sources created in one batch): the batched setters then
// touch one contiguous run of the mirrors.
int contiguous_run(sbm_engine* e, uint32_t n, const int* ids) {
	if (n < 1024) return -1;
	const int s0 = slot_for(e, ids[0]);
	if (s0 < 0 || (size_t)ids[0] + n > e->slot_of_id.size()) return -1;
	for (uint32_t i = 1; i < n; ++i) if (ids[i] != ids[0] + (int)i) return -1;
	const int* so = e->slot_of_id.data() + ids[0];
	for (uint32_t i = 1; i < n; ++i) if (so[i] != s0 + (int)i) return -1;
	return s0;
}

void apply_setter(sbm_engine* e, SetterRec* r) {
	unsigned char* p = setter_payload(r);
	const uint32_t n = r->n;
	switch (r->kind) {
	case kSetPositions: {
		const int* ids = (const int*)p; const float* xyz = (const float*)(ids + n);
		const int s0 = contiguous_run(e, n, ids);
		for (uint32_t i = 0; i < n; ++i) {
			const int slot = s0 >= 0 ? s0 + (int)i : slot_for(e, ids[i]);
			if (slot < 0) { e->stats.ignored_setter_items++; continue; }
			const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
			// AudioEngine drops non-finite poses before they reach Resonance (AudioEngine.cpp:951).
			if (!(std::isfinite(x) && std::isfinite(y) && std::isfinite(z))) continue;
			F4& q = e->h_posvol[(uint32_t)slot];
			q.x = x; q.y = y; q.z = z;
			if (s0 < 0) e->d_posvol.mark((uint32_t)slot);
		}
		if (s0 >= 0) e->d_posvol.mark_range((uint32_t)s0, (uint32_t)s0 + n);
		break;
	}
	case kSetVolumes: case kSetOcclusions: {
		const int* ids = (const int*)p; const float* v = (const float*)(ids + n);
		const bool vol = r->kind == kSetVolumes;
		DirtySet& ds = vol ? e->d_posvol : e->d_occl;
		const int s0 = contiguous_run(e, n, ids);
		for (uint32_t i = 0; i < n; ++i) {
			const int slot = s0 >= 0 ? s0 + (int)i : slot_for(e, ids[i]);
			if (slot < 0) { e->stats.ignored_setter_items++; continue; }
			if (vol) e->h_posvol[(uint32_t)slot].w = v[i]; else e->h_occl[(uint32_t)slot] = v[i];
			if (s0 < 0) ds.mark((uint32_t)slot);
		}
		if (s0 >= 0) ds.mark_range((uint32_t)s0, (uint32_t)s0 + n);
		break;
	}
	case kSetHead: {
		const float* f = (const float*)p;
		memcpy(e->g.head_pos, f, 12); memcpy(e->g.head_q, f + 3, 16);
		break;
	}
	case kSetMaster: e->g.master = *(const float*)p; break;
	case kSetMixParams: {
		const double* delta = (const double*)p; const int* ids = (const int*)(delta + n); const int* z = ids + n; const float* mixf = (const float*)(z + n);
		for (uint32_t i = 0; i < n; ++i) {
			const int slot = slot_for(e, ids[i]);
			if (slot < 0) { e->stats.ignored_setter_items++; continue; }
			const SrcStatic& st = e->h_stat[(uint32_t)slot];
			if ((uint32_t)z[i] >= st.nz) { e->stats.ignored_setter_items++; continue; }
			MixZ& m = e->h_mixz[st.mix_first + (uint32_t)z[i]];
			m.delta = delta[i]; m.mix_factor = mixf[i];
			e->d_mixz.mark(st.mix_first + (uint32_t)z[i]);
		}
		break;
	}
	default: break;
	}
}

// Applies every queued setter call in the order it was queued (mixer thread, under e->mu).
void drain_setters(sbm_engine* e) {
	SetterRec* l = e->setq.exchange(nullptr, std::memory_order_acquire);
	SetterRec* fifo = nullptr;
	while (l) { SetterRec* nx = l->next; l->next = fifo; fifo = l; l = nx; }     // the stack holds the newest call first
	while (fifo) { SetterRec* nx = fifo->next; apply_setter(e, fifo); free(fifo); fifo = nx; }
}

// ---- the per-buffer upload block ----
// One pinned block = [kMaxUpSegs segment descriptors][values / element indices of every segment][patch lists], one H2D,
// one kernel (k_apply_uploads) that copies or scatters every segment into its device array.
struct SegPlan { void* dst; const unsigned char* src; DirtySet* ds; uint32_t elem_bytes; bool scatter; uint32_t count, lo; };

template <class T> void plan_dirty(std::vector<SegPlan>& plans, T* dv, const std::vector<T>& mirror, DirtySet& ds) {
	if (!ds.any()) return;
	SegPlan p;
	p.dst = dv; p.src = (const unsigned char*)mirror.data(); p.ds = &ds; p.elem_bytes = (uint32_t)sizeof(T);
	const size_t span = (size_t)(ds.hi - ds.lo);
	p.scatter = !ds.dense && ds.idx.size() * (sizeof(T) + 4) < span * sizeof(T);
	p.count = p.scatter ? (uint32_t)ds.idx.size() : (uint32_t)span;
	p.lo = ds.lo;
	plans.push_back(p);
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
	// K1 beside K2 (mix_begin_locked) is opt-in: measured on B200 at 2^20 sources it is SLOWER than one kernel after the other
	// (0.56 vs 0.54 ms per buffer, DESIGN.md "K1 beside K2") -- kept, tested and selectable for the record and for further work.
	if (const char* v = getenv("SBM_OVERLAP")) e->overlap_allowed = v[0] && v[0] != '0';
	if (const char* v = getenv("SBM_OVERLAP_MODE")) e->overlap_mode = atoi(v);      // development: 2 = persistent K1, mix kernel waits for all of it
	e->cfg = *cfg;
	e->cfg.sh_hrir = nullptr;   // caller's pointer is not retained
	memset(&e->stats, 0, sizeof(e->stats));
	if (dev::init(cfg->device, &e->stream)) { const int rc = SBM_ERR_CUDA; fprintf(stderr, "sbm_create: %s\n", dev::last_error()); delete e; return rc; }
	e->upload_stream = dev::stream_create();
	e->copy_stream = dev::stream_create();
	e->early_stream = dev::stream_create();
	e->ev_k1_done = dev::event_create();
	e->ev_early = dev::event_create();
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
	e->slot_buf.assign(cap, -1);
	e->mixz_buf.assign(mixcap, -1);
	e->d_stat.init(cap); e->d_posvol.init(cap); e->d_occl.init(cap); e->d_ring.init(cap); e->d_mixz.init(mixcap);
	bool ok = true;
	SlotArrays& A = e->A;
	ok = ok && dalloc(e->dv_stat, cap) && dalloc(e->dv_posvol, cap) && dalloc(e->dv_occl, cap) && dalloc(e->dv_ring_write, cap) && dalloc(e->dv_mixz, mixcap);
	A.stat = e->dv_stat; A.posvol = e->dv_posvol; A.occl = e->dv_occl; A.ring_write = e->dv_ring_write; A.mixz = e->dv_mixz;
	ok = ok && dalloc(A.live, cap) && dalloc(A.read_i, cap) && dalloc(A.eof, cap) && dalloc(A.rs, cap) && dalloc(A.dgain, cap) &&
	     dalloc(A.occ_cur, cap) && dalloc(A.lpf_prev, cap) && dalloc(A.bits, cap) && dalloc(A.rec, cap) && dalloc(A.mrec, cap) && dalloc(A.rfl, cap) && dalloc(A.rgain, cap) && dalloc(A.ready, (size_t)cap / kK1Tile + 2) && dalloc(A.finished, 2 * (size_t)cap) &&
	     dalloc(A.mix_pos, mixcap) && dalloc(A.mix_pos0, mixcap);
	A.cstride = (int)cap;      // cgain is quad-major: [kNumSh / 4][cap] x 16 bytes
	ok = ok && dalloc(A.cgain, (size_t)cap * kNumSh);
	const bool hrtf = cfg->sh_hrir != nullptr;
	if (hrtf) ok = ok && dalloc(A.shg, cap) && dalloc(A.shr, cap); else { A.shg = nullptr; A.shr = nullptr; }
	ok = ok && dalloc(e->dv_flags, 1) && dalloc(e->dv_table, (size_t)kTableAz * kTableEl * kShRaw);
	ok = ok && dalloc(e->dv_lists, cap) && dalloc(e->dv_list_ring, cap) && dalloc(e->dv_list_mix, cap);
	A.list_mix = e->dv_list_mix;
	e->bus_floats = (size_t)kBusRows * F + 4;
	ok = ok && dalloc(e->dv_bus, e->bus_floats);
	e->partials_floats = dev::partials_floats(F);
	ok = ok && dalloc(e->dv_partials, e->partials_floats);
	e->out_bytes = sizeof(float) * 2 * F + sizeof(OutTail) + sizeof(unsigned long long) * (size_t)cap + sizeof(int) * kInlineFinished;
	e->out_bytes = (e->out_bytes + 255) & ~(size_t)255;
	e->dv_out = (float*)dev::alloc(2 * e->out_bytes);
	for (int k = 0; k < 2; ++k) {
		e->fl[k].h_out = (unsigned char*)dev::alloc_pinned(e->out_bytes);
		e->fl[k].dv_finished = A.finished ? A.finished + (size_t)k * cap : nullptr;
		ok = ok && e->fl[k].h_out;
	}
	ok = ok && e->dv_out && e->upload_stream && e->copy_stream && e->early_stream;
	if (ok && hrtf) {
		const size_t T = (size_t)cfg->sh_hrir_taps;
		ok = dalloc(e->dv_hrir, kNumSh * T) && dalloc(e->dv_hist, kNumSh * T) && dalloc(e->dv_dec, 2 * (size_t)F);
		if (ok) ok = dev::h2d(e->dv_hrir, cfg->sh_hrir, sizeof(float) * kNumSh * T, e->stream) == 0;
	}
	if (ok) ok = dev::h2d(e->dv_table, encoder_table(), sizeof(float) * kTableAz * kTableEl * kShRaw, e->stream) == 0;
	if (ok) {
		// AmbisonicBinauralDecoderNode starts with num_frames_processed_on_empty_input_ = frames (no flush pending).
		DevFlags f; memset(&f, 0, sizeof(f)); f.sh_flush_frames = (int)F; f.sh1_flush_frames = (int)F;
		ok = dev::h2d(e->dv_flags, &f, sizeof(f), e->stream) == 0 && dev::sync(e->stream) == 0;
	}
	if (!ok) { fprintf(stderr, "sbm_create: %s\n", dev::last_error()); sbm_destroy(e); return SBM_ERR_CUDA; }
	memset(&e->g, 0, sizeof(e->g));
	e->g.head_q[3] = 1.0f; e->g.master = 1.0f; e->g.frames = F; e->g.rate = cfg->sample_rate_hz;
	for (int k = 0; k < 2; ++k) {
		e->fl[k].done = dev::event_create();
		e->fl[k].k4_done = dev::event_create();
		for (int i = 0; i < 4; ++i) e->fl[k].ev[i] = dev::event_create();
	}
	*out_engine = e;
	return SBM_OK;
}

SBM_API void sbm_destroy(sbm_engine* e) {
	if (!e) return;
	dev::use(e->cfg.device);
	if (e->stream) dev::sync(e->stream);
	if (e->upload_stream) dev::sync(e->upload_stream);
	if (e->copy_stream) dev::sync(e->copy_stream);
	if (e->early_stream) dev::sync(e->early_stream);
	for (SetterRec* l = e->setq.exchange(nullptr); l;) { SetterRec* nx = l->next; free(l); l = nx; }
	SlotArrays& A = e->A;
	void* ptrs[] = {e->dv_stat, e->dv_posvol, e->dv_occl, e->dv_ring_write, e->dv_mixz, A.live, A.read_i, A.eof, A.rs, A.dgain, A.occ_cur,
	                A.lpf_prev, A.bits, A.rec, A.mrec, A.rfl, A.rgain, A.ready, e->room.dv_params, e->room.dv_state, e->room.dv_delay, e->room.dv_foa, e->room.dv_hrir1,
	                e->room.dv_hist1, e->room.dv_dec1, e->fl[0].dv_finished, A.mix_pos, A.mix_pos0, A.cgain, A.shg, A.shr, e->dv_flags,
	                e->dv_table, e->dv_lists, e->dv_list_ring, e->dv_list_mix, e->dv_bus, e->dv_partials, e->dv_out, e->dv_hrir,
	                e->dv_hist, e->dv_dec};
	for (void* p : ptrs) if (p) dev::release(p);
	for (auto& b : e->buffers) if (b.owned && b.dptr) dev::release(b.dptr);
	for (auto& r : e->rings) if (r.storage) dev::release(r.storage);
	for (int k = 0; k < 2; ++k) {
		if (e->fl[k].h_out) dev::release_pinned(e->fl[k].h_out);
		if (e->fl[k].done) dev::event_destroy(e->fl[k].done);
		if (e->fl[k].k4_done) dev::event_destroy(e->fl[k].k4_done);
		for (int i = 0; i < 4; ++i) if (e->fl[k].ev[i]) dev::event_destroy(e->fl[k].ev[i]);
	}
	for (void* m : e->peer_mapped) dev::peer_close(m);
	if (e->dv_gather) dev::release(e->dv_gather);
	if (e->dv_bus_sum) dev::release(e->dv_bus_sum);
	free_retired(e, true);
	if (e->stage.h) dev::release_pinned(e->stage.h);
	if (e->stage.d) dev::release(e->stage.d);
	if (e->upload_stream) dev::shutdown(e->upload_stream);
	if (e->copy_stream) dev::shutdown(e->copy_stream);
	if (e->early_stream) dev::shutdown(e->early_stream);
	if (e->ev_k1_done) dev::event_destroy(e->ev_k1_done);
	if (e->ev_early) dev::event_destroy(e->ev_early);
	dev::shutdown(e->stream);
	delete e;
}

SBM_API const char* sbm_last_error(const sbm_engine* e) { return e ? e->error.c_str() : "null engine"; }

// ---------------------------------------------------------------------------------------------------------------
// PCM buffers
// ---------------------------------------------------------------------------------------------------------------
SBM_API int sbm_buffer_create(sbm_engine* e, const float* samples, size_t n, int* out_buffer) {
	if (!e || !samples || n == 0 || !out_buffer || n > 0xfffffff0ull) return SBM_ERR_INVALID_ARG;
	dev::use(e->cfg.device);
	// allocation + upload happen outside the engine lock, on their own stream: a long sound file neither waits for the
	// buffers in flight nor keeps the mixer thread out of the engine
	float* d = (float*)dev::alloc(sizeof(float) * (n + 8));   // slack: the mix kernel may read whole 32-byte sectors
	if (!d) return fail_cuda(e, "sbm_buffer_create");
	if (dev::h2d(d, samples, sizeof(float) * n, e->upload_stream) || dev::sync(e->upload_stream)) { dev::release(d); return fail_cuda(e, "sbm_buffer_create"); }
	std::lock_guard<std::mutex> lk(e->mu);
	e->buffers.push_back(BufferInfo{d, n, true, 0, false});
	*out_buffer = (int)e->buffers.size() - 1;
	return SBM_OK;
}

SBM_API int sbm_buffer_adopt_device(sbm_engine* e, const float* device_base, size_t n, size_t stride_samples, int count, int* out_first_buffer) {
	if (!e || !device_base || n == 0 || count <= 0 || !out_first_buffer || n > 0xfffffff0ull) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	*out_first_buffer = (int)e->buffers.size();
	e->buffers.reserve(e->buffers.size() + (size_t)count);
	for (int k = 0; k < count; ++k) e->buffers.push_back(BufferInfo{const_cast<float*>(device_base) + (size_t)k * stride_samples, n, false, 0, false});
	return SBM_OK;
}

SBM_API int sbm_buffer_destroy(sbm_engine* e, int buffer) {
	if (!e) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	if (buffer < 0 || (size_t)buffer >= e->buffers.size() || !e->buffers[(size_t)buffer].dptr) return SBM_ERR_UNKNOWN_ID;
	BufferInfo& b = e->buffers[(size_t)buffer];
	b.doomed = true;
	if (!b.refs) { if (b.owned) retire(e, nullptr, b.dptr); b.dptr = nullptr; b.n = 0; }
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
		if (buffers[i] < 0 || (size_t)buffers[i] >= e->buffers.size() || !e->buffers[(size_t)buffers[i]].dptr || e->buffers[(size_t)buffers[i]].doomed) { rc = SBM_ERR_INVALID_ARG; continue; }
		const int rate = sampling_rate ? sampling_rate[i] : e->cfg.sample_rate_hz;
		// AudioEngine::addSource rejects rates outside 8000..48000 (AudioEngine.cpp:884-885).
		if (rate < 8000 || rate > 48000) { rc = fail(e, SBM_ERR_INVALID_ARG, "Unsupported sampling rate for audio source"); continue; }
		detach_slot_refs(e, (uint32_t)slot);
		BufferInfo& b = e->buffers[(size_t)buffers[i]];
		b.refs++; e->slot_buf[(uint32_t)slot] = buffers[i];
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
	dev::use(e->cfg.device);
	float* storage = (float*)dev::alloc(sizeof(float) * ((size_t)capacity_samples + 8));
	if (!storage) return fail_cuda(e, "sbm_source_attach_ring");
	std::lock_guard<std::mutex> lk(e->mu);
	const int slot = slot_for(e, id);
	if (slot < 0) { dev::release(storage); return SBM_ERR_UNKNOWN_ID; }
	detach_slot_refs(e, (uint32_t)slot);
	RingInfo& r = e->rings[(uint32_t)slot];
	if (r.storage) retire(e, nullptr, r.storage);
	r = RingInfo{storage, (uint32_t)capacity_samples, 0, 0};
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

// Batched ring push: the packets of MANY ring-fed sources in one call (what the voice / streaming threads deliver between
// two buffers).  Packet table and payloads are staged in the engine's pinned staging ring, cross PCIe as ONE copy, and a
// scatter kernel writes each packet into its source's ring -- no stream synchronisation.  offsets[k] .. offsets[k + 1] is
// packet k inside `samples`.  Either every packet is accepted or none (ring full / unknown id -> nothing is written).
static int ring_push_locked(sbm_engine* e, int n, const int* ids, const float* samples, const size_t* offsets) {
	const size_t total = offsets[n] - offsets[0];
	std::vector<RingPush> pk((size_t)n);
	std::vector<uint32_t> slots((size_t)n);
	std::map<uint32_t, unsigned long long> pending;     // samples accepted earlier in this call, per slot
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
	size_t off;
	if (stage_alloc(e, bytes, &off)) return fail_cuda(e, "sbm_source_ring_push(staging)");
	unsigned char* hs = e->stage.h + off;
	unsigned char* ds = e->stage.d + off;
	memcpy(hs, pk.data(), sizeof(RingPush) * (size_t)n);
	// A payload that already lives in page-locked memory (sbm_host_alloc_pinned, cudaHostRegister ...) is copied to the
	// device straight from the caller's buffer (see the header for how long it must stay unmodified); anything else is
	// first copied into the staging ring.
	const bool direct = total && dev::is_pinned(samples + offsets[0]);
	int rc;
	if (direct) {
		rc = dev::h2d(ds, hs, table_bytes, e->stream) || dev::h2d(ds + table_bytes, samples + offsets[0], sizeof(float) * total, e->stream);
	} else {
		memcpy(hs + table_bytes, samples + offsets[0], sizeof(float) * total);
		rc = dev::h2d(ds, hs, bytes, e->stream);
	}
	if (rc || dev::launch_ring_scatter((const RingPush*)ds, (uint32_t)n, (const float*)(ds + table_bytes), e->stream))
		return fail_cuda(e, "sbm_source_ring_push");
	e->pending_ring_h2d_bytes += bytes;
	for (int k = 0; k < n; ++k) {
		RingInfo& r = e->rings[slots[(size_t)k]];
		r.written += pk[(size_t)k].n;
		e->h_ring_write[slots[(size_t)k]] = r.written;
		e->d_ring.mark(slots[(size_t)k]);
	}
	return SBM_OK;
}

SBM_API int sbm_source_ring_push_batch(sbm_engine* e, int n, const int* ids, const float* samples, const size_t* offsets) {
	if (!e || n < 0 || (n && (!ids || !samples || !offsets))) return SBM_ERR_INVALID_ARG;
	if (!n) return SBM_OK;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	return ring_push_locked(e, n, ids, samples, offsets);
}

SBM_API int sbm_source_ring_push(sbm_engine* e, int id, const float* samples, size_t n) {
	if (!e || (!samples && n)) return SBM_ERR_INVALID_ARG;
	if (!n) return SBM_OK;
	const size_t offsets[2] = {0, n};
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	return ring_push_locked(e, 1, &id, samples, offsets);
}

// Page-locked host memory for callers that do not link CUDA themselves (packet payloads for sbm_source_ring_push_batch,
// PCM for sbm_buffer_create): copies from it need no staging.
SBM_API void* sbm_host_alloc_pinned(size_t bytes) { return dev::alloc_pinned(bytes); }
SBM_API void sbm_host_free_pinned(void* p) { dev::release_pinned(p); }

SBM_API int sbm_source_ring_size(sbm_engine* e, int id, uint64_t* out_size) {
	if (!e || !out_size) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
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
	for (int z = 0; z < nz; ++z)
		if (buffers[z] < 0 || (size_t)buffers[z] >= e->buffers.size() || !e->buffers[(size_t)buffers[z]].dptr || e->buffers[(size_t)buffers[z]].doomed) return SBM_ERR_INVALID_ARG;
	SrcStatic& st = e->h_stat[(uint32_t)slot];
	// a range of nz MixZ entries: the slot's own when it already has one of that length, else a recycled one, else fresh
	uint32_t first;
	std::vector<uint32_t>* fr = nullptr;
	if (st.nz != (uint32_t)nz) { auto it = e->mix_free.find((uint32_t)nz); if (it != e->mix_free.end() && !it->second.empty()) fr = &it->second; }
	if (st.nz != (uint32_t)nz && !fr && e->mix_used + (uint32_t)nz > e->h_mixz.size()) return fail(e, SBM_ERR_CAPACITY, "max_mix_sources exceeded");
	if (st.nz == (uint32_t)nz) {
		first = st.mix_first;
		for (int z = 0; z < nz; ++z) { buffer_unref(e, e->mixz_buf[first + (uint32_t)z]); e->mixz_buf[first + (uint32_t)z] = -1; }
		if (e->slot_buf[(uint32_t)slot] >= 0) { buffer_unref(e, e->slot_buf[(uint32_t)slot]); e->slot_buf[(uint32_t)slot] = -1; }
	} else {
		detach_slot_refs(e, (uint32_t)slot);
		if (fr) { first = fr->back(); fr->pop_back(); }
		else { first = e->mix_used; e->mix_used += (uint32_t)nz; }
	}
	for (int z = 0; z < nz; ++z) {
		BufferInfo& b = e->buffers[(size_t)buffers[z]];
		b.refs++; e->mixz_buf[first + (uint32_t)z] = buffers[z];
		e->h_mixz[first + (uint32_t)z] = MixZ{b.dptr, (uint32_t)b.n, mixf[z], delta[z], 0.0};
		e->d_mixz.mark(first + (uint32_t)z);
		e->mix_patches.push_back(MixPatch{first + (uint32_t)z, 0, pos[z]});
	}
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

SBM_API int sbm_source_set_sampling_rate(sbm_engine* e, int id, int sampling_rate) {
	if (!e) return SBM_ERR_INVALID_ARG;
	if (sampling_rate < 8000 || sampling_rate > 48000) return fail(e, SBM_ERR_INVALID_ARG, "Unsupported sampling rate for audio source");
	std::lock_guard<std::mutex> lk(e->mu);
	const int slot = slot_for(e, id);
	if (slot < 0) return SBM_ERR_UNKNOWN_ID;
	SrcStatic& st = e->h_stat[(uint32_t)slot];
	if (((st.flags >> kFlagFeedShift) & 3) == kFeedMix) return fail(e, SBM_ERR_UNSUPPORTED, "mix_sources at a rate other than the engine rate");
	st.rate = sampling_rate;
	e->d_stat.mark((uint32_t)slot);
	add_patch(e, SlotPatch{(uint32_t)slot, kPatchResampler, 0ull, 0, sampling_rate, 0});
	return SBM_OK;
}

SBM_API int sbm_source_set_eof_marker(sbm_engine* e, int id, int64_t samples_to_eof) {
	if (!e) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	const int slot = slot_for(e, id);
	if (slot < 0) return SBM_ERR_UNKNOWN_ID;
	add_patch(e, SlotPatch{(uint32_t)slot, kPatchEof, 0ull, (long long)samples_to_eof, 0, 0});
	return SBM_OK;
}

SBM_API int sbm_source_set_looping(sbm_engine* e, int id, int looping) {
	if (!e) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	const int slot = slot_for(e, id);
	if (slot < 0) return SBM_ERR_UNKNOWN_ID;
	SrcStatic& st = e->h_stat[(uint32_t)slot];
	if (looping) st.flags |= kFlagLooping; else st.flags &= ~kFlagLooping;
	e->d_stat.mark((uint32_t)slot);
	return SBM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// setters: lock-free, any thread, never wait for the mixer or the GPU.  Each call is one record in an MPSC queue, applied by
// the mixer thread at the start of the next buffer in the order queued.  Ids that are unknown (or already destroyed) by
// then are ignored like Resonance ignores them (resonance_audio_api_impl.cc:551-554) and counted in sbm_stats.
// ---------------------------------------------------------------------------------------------------------------
static int queue_ids_floats(sbm_engine* e, uint32_t kind, int n, const int* ids, const float* v, int per) {
	if (!e || n < 0 || (n && (!ids || !v))) return SBM_ERR_INVALID_ARG;
	if (!n) return SBM_OK;
	SetterRec* r = setter_new(kind, (uint32_t)n, (size_t)n * (4 + 4 * (size_t)per));
	if (!r) return SBM_ERR_NO_MEMORY;
	unsigned char* p = setter_payload(r);
	memcpy(p, ids, 4 * (size_t)n);
	memcpy(p + 4 * (size_t)n, v, 4 * (size_t)per * (size_t)n);
	setter_push(e, r);
	return SBM_OK;
}
SBM_API int sbm_set_source_positions(sbm_engine* e, int n, const int* ids, const float* xyz) { return queue_ids_floats(e, kSetPositions, n, ids, xyz, 3); }
SBM_API int sbm_set_source_volumes(sbm_engine* e, int n, const int* ids, const float* volume) { return queue_ids_floats(e, kSetVolumes, n, ids, volume, 1); }
SBM_API int sbm_set_source_occlusions(sbm_engine* e, int n, const int* ids, const float* occ) { return queue_ids_floats(e, kSetOcclusions, n, ids, occ, 1); }

SBM_API int sbm_source_set_mix_params(sbm_engine* e, int n, const int* ids, const int* z, const double* source_delta, const float* mix_factor) {
	if (!e || n < 0 || !ids || !source_delta || !mix_factor) return SBM_ERR_INVALID_ARG;
	if (!n) return SBM_OK;
	SetterRec* r = setter_new(kSetMixParams, (uint32_t)n, (size_t)n * 20);
	if (!r) return SBM_ERR_NO_MEMORY;
	double* d = (double*)setter_payload(r); int* pi = (int*)(d + n); int* pz = pi + n; float* pm = (float*)(pz + n);
	memcpy(d, source_delta, 8 * (size_t)n); memcpy(pi, ids, 4 * (size_t)n);
	if (z) memcpy(pz, z, 4 * (size_t)n); else memset(pz, 0, 4 * (size_t)n);
	memcpy(pm, mix_factor, 4 * (size_t)n);
	setter_push(e, r);
	return SBM_OK;
}

SBM_API int sbm_set_head(sbm_engine* e, const float* pos, const float* q) {
	if (!e || !pos || !q) return SBM_ERR_INVALID_ARG;
	for (int i = 0; i < 3; ++i) if (!std::isfinite(pos[i])) return SBM_OK;   // AudioEngine.cpp:985
	for (int i = 0; i < 4; ++i) if (!std::isfinite(q[i])) return SBM_OK;
	SetterRec* r = setter_new(kSetHead, 0, 28);
	if (!r) return SBM_ERR_NO_MEMORY;
	memcpy(setter_payload(r), pos, 12); memcpy(setter_payload(r) + 12, q, 16);
	setter_push(e, r);
	return SBM_OK;
}

SBM_API int sbm_set_master_volume(sbm_engine* e, float volume) {
	if (!e) return SBM_ERR_INVALID_ARG;
	SetterRec* r = setter_new(kSetMaster, 0, 4);
	if (!r) return SBM_ERR_NO_MEMORY;
	memcpy(setter_payload(r), &volume, 4);
	setter_push(e, r);
	return SBM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// room effects: early reflections of a shoebox room (what Substrata turns on inside every parcel, AudioEngine.cpp:738-768)
// ---------------------------------------------------------------------------------------------------------------
SBM_API int sbm_set_room_sh_hrir(sbm_engine* e, const float* sh_hrir_order1, int taps) {
	if (!e || !sh_hrir_order1 || taps < 1 || taps > 1024) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	sbm_engine::Room& r = e->room;
	if (r.configured) return fail(e, SBM_ERR_INVALID_ARG, "room SH-HRIRs already set");
	if (e->begun || e->collected != e->buffer_index) return fail(e, SBM_ERR_INVALID_ARG, "sbm_set_room_sh_hrir with a buffer in flight");
	const uint32_t F = e->g.frames;
	r.taps1 = (uint32_t)taps;
	r.delay_size = F + 2u * (uint32_t)e->cfg.sample_rate_hz;     // DelayFilter(max_delay = kMaxDelayTimeSeconds * rate, frames): reflections_processor.cc:54,67
	bool ok = dalloc(r.dv_params, 2) && dalloc(r.dv_state, 1) && dalloc(r.dv_delay, r.delay_size) && dalloc(r.dv_foa, 4 * (size_t)F) &&
	          dalloc(r.dv_hrir1, 4 * (size_t)taps) && dalloc(r.dv_hist1, 4 * (size_t)taps) && dalloc(r.dv_dec1, 2 * (size_t)F);
	if (ok) {
		RoomState st; memset(&st, 0, sizeof(st));
		st.frames_processed_on_empty = F;      // reflections_node.cc:31-32
		st.rot_cur[3] = 1.0f;
		ok = dev::h2d(r.dv_hrir1, sh_hrir_order1, sizeof(float) * 4 * (size_t)taps, e->stream) == 0 && dev::h2d(r.dv_state, &st, sizeof(st), e->stream) == 0 &&
		     dev::sync(e->stream) == 0;
	}
	if (!ok) return fail_cuda(e, "sbm_set_room_sh_hrir");
	r.configured = true;
	return SBM_OK;
}

SBM_API int sbm_enable_room_effects(sbm_engine* e, int enable) {
	if (!e) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	if (enable && !e->room.configured) return fail(e, SBM_ERR_INVALID_ARG, "room effects need the order-1 SH-HRIRs first (sbm_set_room_sh_hrir)");
	if (e->room.enabled && !enable) {
		// the sends are muted from the next buffer on; the ReflectionsNode plays out its delay line (2 s at most) on silence
		e->room.tail_buffers = (2u * (uint32_t)e->cfg.sample_rate_hz + 2u * e->g.frames) / e->g.frames + 4u;
	}
	e->room.enabled = enable != 0;
	return SBM_OK;
}

SBM_API int sbm_set_reflection_properties(sbm_engine* e, const float* room_position, const float* room_rotation_xyzw, const float* room_dimensions,
                                          const float* coefficients6, float cutoff_frequency, float gain) {
	if (!e || !room_position || !room_rotation_xyzw || !room_dimensions || !coefficients6) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	sbm_engine::Room& r = e->room;
	memcpy(r.pos, room_position, 12); memcpy(r.rot, room_rotation_xyzw, 16); memcpy(r.dims, room_dimensions, 12); memcpy(r.coeffs, coefficients6, 24);
	r.cutoff = cutoff_frequency; r.gain = gain;
	return SBM_OK;
}

// What ResonanceAudioApiImpl::ProcessNextBuffer does for the room before the graph runs (resonance_audio_api_impl.cc:464-467):
// ReflectionsNode::Update (reflections_node.cc:41-82) compares the properties and the listener position with its cached copies and,
// on any change, has the ReflectionsProcessor recompute the six reflections (reflections_processor.cc:74-97,
// RA/dsp/shoe_box_room.cc:23-61) and start a crossfade.  Float for float like the reference (this file is built without FMA
// contraction); the result rides to the device as this buffer's RoomParams.
static bool floats_differ(const float* a, const float* b, int n) { for (int i = 0; i < n; ++i) if (a[i] != b[i]) return true; return false; }
static void room_update_host(sbm_engine* e) {
	sbm_engine::Room& r = e->room;
	RoomParams& p = r.h_params;
	memset(&p, 0, sizeof(p));
	p.enabled = r.enabled ? 1u : 0u;
	p.delay_size = r.delay_size;
	p.head_q_inv[0] = -e->g.head_q[0]; p.head_q_inv[1] = -e->g.head_q[1]; p.head_q_inv[2] = -e->g.head_q[2]; p.head_q_inv[3] = e->g.head_q[3];
	if (!r.enabled) return;
	const float* listener = e->g.head_pos;
	const bool changed = floats_differ(r.n_pos, r.pos, 3) || floats_differ(r.n_rot, r.rot, 4) || floats_differ(r.n_dims, r.dims, 3) || r.n_cutoff != r.cutoff ||
	                     floats_differ(r.n_coeffs, r.coeffs, 6) || floats_differ(r.n_listener, listener, 3);
	if (!changed) return;
	memcpy(r.n_pos, r.pos, 12); memcpy(r.n_rot, r.rot, 16); memcpy(r.n_dims, r.dims, 12); memcpy(r.n_coeffs, r.coeffs, 24);
	r.n_cutoff = r.cutoff; memcpy(r.n_listener, listener, 12);
	const float rate = (float)e->cfg.sample_rate_hz;
	// ComputeLowPassMonoPoleCoefficient (RA/dsp/filter_coefficient_generators.cc:151-163) + MonoPoleFilter::SetCoefficient's clamp
	float coef = 0.0f;
	if (r.n_cutoff > 20.0f) { const float itc = 6.28318530717958647692f * r.n_cutoff; coef = rate / (itc + rate); }
	p.lp_coef = std::max(std::min(coef, 1.0f), 0.0f);
	// the listener in the room's frame: GetRelativeDirection(room position, conj(room rotation), listener)
	const float conj[4] = {-r.n_rot[0], -r.n_rot[1], -r.n_rot[2], r.n_rot[3]};
	float rel[3];
	source_relative(r.n_pos, conj, r.n_listener, rel);
	const bool inside = fabsf(rel[0] - 0.0f) <= 0.5f * r.n_dims[0] && fabsf(rel[1] - 0.0f) <= 0.5f * r.n_dims[1] && fabsf(rel[2] - 0.0f) <= 0.5f * r.n_dims[2];
	float delay_s[6] = {0, 0, 0, 0, 0, 0}, mag[6] = {0, 0, 0, 0, 0, 0}, max_delay = 0.0f;
	if (inside) {
		const float off[3] = {0.5f * r.n_dims[0], 0.5f * r.n_dims[1], 0.5f * r.n_dims[2]};
		const float dist[6] = {off[0] + rel[0] + 1.0f, off[0] - rel[0] + 1.0f, off[1] + rel[1] + 1.0f, off[1] - rel[1] + 1.0f, off[2] + rel[2] + 1.0f, off[2] - rel[2] + 1.0f};
		for (int i = 0; i < 6; ++i) { delay_s[i] = dist[i] / 343.0f; mag[i] = r.n_coeffs[i] / dist[i]; }
	}
	const size_t max_delay_samples = 2u * (size_t)e->cfg.sample_rate_hz;
	for (int i = 0; i < 6; ++i) {
		max_delay = std::max(max_delay, delay_s[i]);
		p.tgt_delay[i] = (uint32_t)std::min(max_delay_samples, (size_t)(delay_s[i] * rate));      // UpdateGainsAndDelays (reflections_processor.cc:123-132)
		p.tgt_gain[i] = mag[i];
	}
	p.frames_to_process_on_empty = (uint32_t)((size_t)e->g.frames + (size_t)(max_delay * rate));
	p.updated = 1;
}

// Scenes from this size on are about throughput: their result goes home on the copy stream (mix_enqueue_finish_locked) and their per-frame
// uploads travel beside the previous buffer's mix kernel (upload_block_locked); smaller ones are about latency and keep everything in order
// in the mixer's stream.
constexpr size_t kSideCopyMinSources = 65536;

// ---------------------------------------------------------------------------------------------------------------
// the per-buffer call
// ---------------------------------------------------------------------------------------------------------------
// One buffer = sbm_mix_begin (queued setters applied, ONE upload block, K1, K2, K3) -> [the caller may sum buses across
// ranks here] -> the finish part (gather, K4, D2H of the result into pinned memory, an event) -> collection (wait for the
// event WITHOUT the engine lock, hand the samples out, host bookkeeping).  sbm_mix / sbm_mix_finish do all of it
// synchronously; sbm_mix_submit + sbm_mix_collect split it so that up to two buffers are in flight and the host's launch
// latency hides behind the previous buffer's kernels -- the mixer thread of the reference also runs ahead of the device
// callback (AudioEngine.cpp:540-545).
static int upload_block_locked(sbm_engine* e, const SlotPatch** d_patches, const MixPatch** d_mix_patches) {
	NvtxRange nvtx("sbm: upload block");
	std::vector<SegPlan> plans;
	// (what only K1 -- positions, volumes, occlusions -- or only this buffer's room kernel -- its parity's RoomParams -- reads)
	const bool only_per_frame = !e->d_stat.any() && !e->d_ring.any() && !e->d_mixz.any() && !e->lists_dirty && (e->d_posvol.any() || e->d_occl.any());
	plan_dirty(plans, e->dv_stat, e->h_stat, e->d_stat);
	plan_dirty(plans, e->dv_posvol, e->h_posvol, e->d_posvol);
	plan_dirty(plans, e->dv_occl, e->h_occl, e->d_occl);
	plan_dirty(plans, e->dv_ring_write, e->h_ring_write, e->d_ring);
	plan_dirty(plans, e->dv_mixz, e->h_mixz, e->d_mixz);
	if (e->lists_dirty) {
		rebuild_lists(e);
		// the render lists travel as plain contiguous segments
		if (!e->list_pan.empty()) plans.push_back(SegPlan{e->dv_lists, (const unsigned char*)e->list_pan.data(), nullptr, 4, false, (uint32_t)e->list_pan.size(), 0});
		if (!e->list_sh.empty()) plans.push_back(SegPlan{e->dv_lists + e->list_pan.size(), (const unsigned char*)e->list_sh.data(), nullptr, 4, false, (uint32_t)e->list_sh.size(), 0});
		if (!e->list_mix_pan.empty()) plans.push_back(SegPlan{e->dv_list_mix, (const unsigned char*)e->list_mix_pan.data(), nullptr, 4, false, (uint32_t)e->list_mix_pan.size(), 0});
		if (!e->list_mix_sh.empty()) plans.push_back(SegPlan{e->dv_list_mix + e->list_pan.size(), (const unsigned char*)e->list_mix_sh.data(), nullptr, 4, false, (uint32_t)e->list_mix_sh.size(), 0});
		if (!e->list_ring.empty()) plans.push_back(SegPlan{e->dv_list_ring, (const unsigned char*)e->list_ring.data(), nullptr, 4, false, (uint32_t)e->list_ring.size(), 0});
	}
	if (e->room.run_this_buffer)      // this buffer's RoomParams -> the parity's device copy
		plans.push_back(SegPlan{e->room.dv_params + (e->buffer_index & 1), (const unsigned char*)&e->room.h_params, nullptr, (uint32_t)sizeof(RoomParams), false, 1, 0});
	*d_patches = nullptr; *d_mix_patches = nullptr;
	if (plans.empty() && e->patches.empty() && e->mix_patches.empty()) return 0;
	if (plans.size() > (size_t)kMaxUpSegs) return -1;
	// layout
	size_t bytes = sizeof(UpSeg) * kMaxUpSegs;
	std::vector<size_t> val_off(plans.size()), idx_off(plans.size());
	for (size_t k = 0; k < plans.size(); ++k) {
		const SegPlan& p = plans[k];
		val_off[k] = bytes; bytes += ((size_t)p.count * p.elem_bytes + 15) & ~(size_t)15;
		idx_off[k] = bytes; if (p.scatter) bytes += ((size_t)p.count * 4 + 15) & ~(size_t)15;
	}
	const size_t patch_off = bytes; bytes += (sizeof(SlotPatch) * e->patches.size() + 15) & ~(size_t)15;
	const size_t mpatch_off = bytes; bytes += (sizeof(MixPatch) * e->mix_patches.size() + 15) & ~(size_t)15;
	size_t off;
	if (stage_alloc(e, bytes, &off)) return -1;
	unsigned char* hs = e->stage.h + off;
	unsigned char* ds = e->stage.d + off;
	UpSeg* segs = (UpSeg*)hs;
	uint32_t max_words = 0;
	for (size_t k = 0; k < plans.size(); ++k) {
		const SegPlan& p = plans[k];
		UpSeg& sg = segs[k];
		sg.elem_words = p.elem_bytes / 4; sg.count = p.count; sg.src_word = (uint32_t)(val_off[k] / 4); sg.pad = 0;
		unsigned char* vals = hs + val_off[k];
		if (p.scatter) {
			uint32_t* ix = (uint32_t*)(hs + idx_off[k]);
			for (uint32_t j = 0; j < p.count; ++j) { const uint32_t i = p.ds->idx[j]; ix[j] = i; memcpy(vals + (size_t)j * p.elem_bytes, p.src + (size_t)i * p.elem_bytes, p.elem_bytes); }
			sg.dst = p.dst; sg.idx_word = (uint32_t)(idx_off[k] / 4);
		} else {
			memcpy(vals, p.src + (size_t)p.lo * p.elem_bytes, (size_t)p.count * p.elem_bytes);
			sg.dst = (unsigned char*)p.dst + (size_t)p.lo * p.elem_bytes; sg.idx_word = 0xffffffffu;
		}
		max_words = std::max(max_words, sg.elem_words * sg.count);
		if (p.ds) p.ds->clear();
	}
	if (!e->patches.empty()) memcpy(hs + patch_off, e->patches.data(), sizeof(SlotPatch) * e->patches.size());
	if (!e->mix_patches.empty()) memcpy(hs + mpatch_off, e->mix_patches.data(), sizeof(MixPatch) * e->mix_patches.size());
	// The per-frame case -- nothing but source positions / volumes / occlusions (and the room parameters) changed, the previous buffer still
	// in flight (the caller pipelines): K1 is the only kernel that reads those arrays (the room kernel its own parity's parameters), so the block crosses PCIe and is scattered on a stream of its own as
	// soon as the previous buffer's K1 is done, i.e. beside that buffer's mix kernel, and this buffer's K1 finds it applied.  Measured on
	// B200: 1 M sources with 4 096 scattered movers per buffer 0.036 -> 0.023 ms on top of the resident time, 262 144 ambisonic sources 0.026 -> 0.006.  Everything
	// else (attach / detach, mix parameters, rings, lists, state patches) stays in the mixer's stream, in order.
	// (large scenes only, like the result's copy stream: a small scene pays more for the hop between streams than the upload holds up --
	// C2 end to end 0.072 -> 0.085 ms with it)
	const bool early = only_per_frame && e->patches.empty() && e->mix_patches.empty() && e->k1_event_valid &&
	                   e->fl[(e->buffer_index + 1) & 1].pending && e->list_pan.size() + e->list_sh.size() >= kSideCopyMinSources;
	void* up = early ? e->early_stream : e->stream;
	if (early && dev::stream_wait_event(up, e->ev_k1_done)) return -1;
	if (dev::h2d(ds, hs, bytes, up)) return -1;
	e->stats.last_h2d_bytes += bytes;
	if (!plans.empty() && dev::launch_apply_uploads((const UpSeg*)ds, (uint32_t)plans.size(), (const uint32_t*)ds, max_words, up)) return -1;
	if (early) {
		dev::event_record(e->ev_early, up);
		if (dev::stream_wait_event(e->stream, e->ev_early)) return -1;
	}
	*d_patches = (const SlotPatch*)(ds + patch_off);
	*d_mix_patches = (const MixPatch*)(ds + mpatch_off);
	return 0;
}

static int mix_begin_locked(sbm_engine* e) {
	NvtxRange nvtx("sbm: mix begin (setters, upload, K1-K3)");
	if (e->begun) return fail(e, SBM_ERR_INVALID_ARG, "sbm_mix_begin called twice without sbm_mix_finish");
	sbm_engine::InFlight& fl = e->fl[e->buffer_index & 1];
	if (fl.pending) return fail(e, SBM_ERR_INVALID_ARG, "two buffers already in flight: call sbm_mix_collect first");
	// bytes pushed into rings since the previous buffer belong to this buffer's host->device traffic
	e->stats.last_h2d_bytes = e->pending_ring_h2d_bytes; e->stats.last_d2h_bytes = 0;
	e->pending_ring_h2d_bytes = 0;
	void* st = e->stream;
	// 1. queued setters -> mirrors; everything the host changed since the last buffer -> ONE block, one copy, one kernel
	drain_setters(e);
	// room effects: the ReflectionsNode runs while the sends are on and until its delay line has played out afterwards
	e->room.run_this_buffer = e->room.configured && (e->room.enabled || e->room.tail_buffers > 0);
	if (e->room.run_this_buffer) {
		if (!e->room.enabled) e->room.tail_buffers--;
		room_update_host(e);
	}
	e->g.room = e->room.enabled ? 1u : 0u;
	e->g.refl_gain = e->room.gain;
	const SlotPatch* d_patches; const MixPatch* d_mix_patches;
	if (upload_block_locked(e, &d_patches, &d_mix_patches)) return fail_cuda(e, "sbm_mix_begin(upload)");
	if (!e->patches.empty() || !e->mix_patches.empty()) {
		if (dev::launch_patch(e->A, d_patches, (uint32_t)e->patches.size(), d_mix_patches, (uint32_t)e->mix_patches.size(), st))
			return fail_cuda(e, "k_patch");
		for (const SlotPatch& p : e->patches) e->patch_of_slot[p.slot] = -1;
		e->patches.clear(); e->mix_patches.clear();
	}
	e->g.parity = (uint32_t)(e->buffer_index & 1);
	e->g.n_slots = e->n_slots;
	e->g.n_pan = (uint32_t)e->list_pan.size(); e->g.n_sh = (uint32_t)e->list_sh.size();
	e->A.finished = fl.dv_finished;      // per-parity list: the previous buffer's may not have been collected yet
	// 2. kernels
	// Large stereo-bus scenes: K1 runs as a persistent grid BESIDE the mix kernel, which consumes its records tile by tile (k1_params.cu).
	// (Not with ambisonic sources or room effects -- those mix kernels leave K1 no registers -- nor at 1024 frames, one CTA per SM.)
	const bool overlap = e->overlap_allowed && e->g.n_sh == 0 && !e->room.enabled && e->g.n_pan >= kOverlapMinSources && e->g.frames <= 512;
	e->g.overlap = overlap ? 1u : 0u;
	e->g.epoch = (uint32_t)(e->buffer_index % 0xfffffffeull) + 1u;
	// (an event between the two launches would serialise them: overlapped, the pair is timed as one span and reported as the mix kernel's)
	if (e->profiling) { dev::event_record(fl.ev[0], st); if (overlap && e->overlap_mode != 2) dev::event_record(fl.ev[1], st); }
	if (e->g.n_pan + e->g.n_sh && dev::launch_params(e->g, e->A, e->dv_flags, e->dv_table, e->dv_lists, e->n_slots, st)) return fail_cuda(e, "k1_params");
	if (e->profiling && !(overlap && e->overlap_mode != 2)) dev::event_record(fl.ev[1], st);
	// (what an early upload of the NEXT buffer waits for; not between an overlapped pair, which an event would serialise)
	e->k1_event_valid = !overlap && (size_t)(e->g.n_pan + e->g.n_sh) >= kSideCopyMinSources;     // (only scenes that use early uploads pay for the event)
	if (e->k1_event_valid) dev::event_record(e->ev_k1_done, st);
	MixArgs m;
	m.ready = overlap && e->overlap_mode != 2 ? e->A.ready : nullptr; m.epoch = e->g.epoch;
	m.rfl = e->room.enabled ? e->A.rfl : nullptr;
	m.mrec = e->A.mrec; m.rec = e->A.rec; m.shg = e->A.shg; m.shr = e->A.shr; m.mixz = e->dv_mixz; m.mix_pos0 = e->A.mix_pos0; m.rs = e->A.rs; m.lpf_prev = e->A.lpf_prev;
	m.n_pan = (uint32_t)e->list_pan.size();
	m.n_sh = (uint32_t)e->list_sh.size();
	m.frames = e->g.frames; m.rate = e->g.rate; m.partials = e->dv_partials; m.partials_floats = e->partials_floats;
	m.bus = e->dv_bus; m.flags = e->dv_flags; m.parity = e->g.parity;
	m.ev_after_k2 = e->profiling ? fl.ev[2] : nullptr;
	m.peer = e->peer; m.peer.parity = (uint32_t)(e->peer_buffers & 1); m.bus_floats = e->bus_floats; m.rot_div = 1;
	if (dev::launch_mix(m, st)) return fail_cuda(e, "k2_mix");
	e->begun = true;
	fl.active = (uint32_t)(e->list_pan.size() + e->list_sh.size());
	fl.h2d_bytes = e->stats.last_h2d_bytes;
	return SBM_OK;
}

// Enqueues the finish part of the buffer begun last; no host synchronisation.
static int mix_enqueue_finish_locked(sbm_engine* e) {
	NvtxRange nvtx("sbm: mix finish (exchange, room, K4, D2H)");
	if (!e->begun) return fail(e, SBM_ERR_INVALID_ARG, "sbm_mix_finish without sbm_mix_begin");
	e->begun = false;
	void* st = e->stream;
	const uint32_t F = e->g.frames;
	const uint32_t parity = (uint32_t)(e->buffer_index & 1);
	sbm_engine::InFlight& fl = e->fl[parity];
	if (e->peer.world) {
		// wait for every rank's bus of this buffer (K3 stored them into our gather area) and sum them in rank order
		PeerArgs pa = e->peer; pa.parity = (uint32_t)(e->peer_buffers & 1); pa.timeout_ns = e->peer_timeout_ns;
		const unsigned int target = (unsigned int)(e->peer.world * ((e->peer_buffers >> 1) + 1));
		if (dev::launch_peer_gather(pa, e->dv_bus, e->dv_bus_sum, e->bus_floats, target, e->dv_flags, st)) return fail_cuda(e, "k_peer_gather");
		e->peer_buffers++;
	}
	if (e->room.run_this_buffer) {
		RoomArgs ra;
		ra.bus = e->peer.world ? e->dv_bus_sum : e->dv_bus; ra.params = e->room.dv_params + parity; ra.state = e->room.dv_state;
		ra.delay_line = e->room.dv_delay; ra.foa = e->room.dv_foa; ra.flags = e->dv_flags; ra.frames = F;
		if (dev::launch_room(ra, st)) return fail_cuda(e, "k_room");
	}
	FinishArgs f;
	f.foa = e->room.dv_foa; f.hrir1 = e->room.configured ? e->room.dv_hrir1 : nullptr; f.taps1 = e->room.taps1; f.hist1 = e->room.dv_hist1; f.dec1 = e->room.dv_dec1;
	f.room_enabled = e->room.enabled ? 1u : 0u;
	f.bus = e->peer.world ? e->dv_bus_sum : e->dv_bus; f.hrir = e->dv_hrir; f.taps = (uint32_t)e->cfg.sh_hrir_taps; f.hist = e->dv_hist; f.dec = e->dv_dec;
	f.flags = e->dv_flags; f.parity = parity; f.frames = F; f.out = (float*)((unsigned char*)e->dv_out + (size_t)parity * e->out_bytes);
	f.list_ring = e->dv_list_ring; f.n_ring = (uint32_t)e->list_ring.size(); f.read_i = e->A.read_i; f.finished = fl.dv_finished; f.hist_smem_all = 0;
	if (dev::launch_finish(f, st)) return fail_cuda(e, "k4_finish");
	if (e->profiling) dev::event_record(fl.ev[3], st);
	const size_t bytes = sizeof(float) * 2 * F + sizeof(OutTail) + sizeof(unsigned long long) * e->list_ring.size() + sizeof(int) * kInlineFinished;
	if (fl.active >= kSideCopyMinSources) {
		// A large scene is about throughput: the result goes home on the copy stream, behind K4b, and the mixer's stream is free for the
		// next buffer's K1 at once (this buffer's result block is not written again before the buffer has been collected: two blocks,
		// two buffers in flight).  Measured on B200: 1 M panned sources 0.513 -> 0.496 ms per buffer, 262 144 ambisonic 0.309 -> 0.297,
		// C3 0.195 -> 0.185.
		dev::event_record(fl.k4_done, st);
		if (dev::stream_wait_event(e->copy_stream, fl.k4_done) || dev::d2h(fl.h_out, f.out, bytes, e->copy_stream)) return fail_cuda(e, "sbm_mix_finish(d2h)");
		dev::event_record(fl.done, e->copy_stream);
	} else {
		// A small scene is about latency: the hop to another stream costs more than the copy holds up (C2 end to end 0.070 -> 0.081 ms).
		if (dev::d2h(fl.h_out, f.out, bytes, st)) return fail_cuda(e, "sbm_mix_finish(d2h)");
		dev::event_record(fl.done, st);
	}
	fl.pending = true; fl.seq = e->buffer_index; fl.list_ring = e->list_ring; fl.d2h_bytes = bytes;
	fl.stage_mark = e->stage.head; fl.stage_gen = e->stage.gen;
	e->buffer_index++;
	return SBM_OK;
}

// Waits for the oldest buffer in flight -- with the engine lock RELEASED: source creation / destruction and ring pushes by
// other threads go on meanwhile -- and hands it out.  `lk` holds e->mu on entry and on return.
static int mix_collect_locked(sbm_engine* e, std::unique_lock<std::mutex>& lk, float* out_interleaved, int* out_valid) {
	NvtxRange nvtx("sbm: mix collect (wait + copy out)");
	sbm_engine::InFlight& fl = e->fl[e->collected & 1];
	if (!fl.pending || fl.seq != e->collected) return fail(e, SBM_ERR_INVALID_ARG, "sbm_mix_collect: no buffer in flight");
	void* st = e->stream;
	const uint32_t F = e->g.frames;
	void* done = fl.done;
	lk.unlock();
	const int wait_rc = dev::event_sync(done);     // only the mixer thread submits / collects: `fl` cannot change under us
	lk.lock();
	if (wait_rc) return fail_cuda(e, "sbm_mix_collect");
	fl.pending = false;
	e->collected++;
	if (fl.stage_gen == e->stage.gen) e->stage.tail = fl.stage_mark;
	free_retired(e, false);
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
	// sources stopped by the feeder (AudioEngine.cpp:499-515): the first kInlineFinished ids ride in the result block
	if (tail.n_finished > 0) {
		std::vector<int> ids((size_t)tail.n_finished);
		const int* inl = (const int*)(reads + fl.list_ring.size());
		if (tail.n_finished <= kInlineFinished) memcpy(ids.data(), inl, sizeof(int) * ids.size());
		else {
			// more than fit inline (a mass expiry): fetch the whole list; the only path that waits for the stream here
			if (dev::d2h(ids.data(), fl.dv_finished, sizeof(int) * ids.size(), st) || dev::sync(st)) return fail_cuda(e, "sbm_mix_finish(finished)");
			e->stats.last_d2h_bytes += sizeof(int) * ids.size();
		}
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
	if (tail.status == kTailPipeTimeout) return fail(e, SBM_ERR_CUDA, "the parameter kernel's records did not reach the mix kernel (overlapped K1 / K2): this buffer is not valid");
	if (tail.status == kTailPeerTimeout) return fail(e, SBM_ERR_PEER_TIMEOUT, "a peer rank did not deliver its bus in time: this buffer holds the local sources only (see sbm_peer_resync)");
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
	std::unique_lock<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	const int rc = mix_enqueue_finish_locked(e);
	if (rc) return rc;
	return mix_collect_locked(e, lk, out_interleaved, out_valid);
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
	std::unique_lock<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	return mix_collect_locked(e, lk, out_interleaved, out_valid);
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
	e->peer_buffers = 0;     // the exchange keeps its own buffer count (parity of the gather area, arrival targets): ranks need not agree on buffer_index
	return SBM_OK;
}

SBM_API int sbm_set_peer_timeout_ms(sbm_engine* e, int timeout_ms) {
	if (!e || timeout_ms <= 0) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	e->peer_timeout_ns = (unsigned long long)timeout_ms * 1000000ull;
	return SBM_OK;
}

SBM_API int sbm_peer_resync(sbm_engine* e) {
	if (!e) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	dev::use(e->cfg.device);
	if (!e->peer.world) return fail(e, SBM_ERR_INVALID_ARG, "not connected");
	if (e->begun || e->collected != e->buffer_index) return fail(e, SBM_ERR_INVALID_ARG, "sbm_peer_resync with a buffer in flight");
	// own arrival counters + the K3 election counter back to zero; the exchange count restarts
	if (dev::zero(e->dv_gather + gather_area_floats(e), sizeof(float) * 16, e->stream) || dev::sync(e->stream)) return fail_cuda(e, "sbm_peer_resync");
	e->peer_buffers = 0;
	return SBM_OK;
}

SBM_API int sbm_bus_device(sbm_engine* e, float** out_device_ptr, size_t* out_num_floats) {
	if (!e || !out_device_ptr || !out_num_floats) return SBM_ERR_INVALID_ARG;
	*out_device_ptr = e->dv_bus; *out_num_floats = e->bus_floats;
	return SBM_OK;
}
SBM_API int sbm_out_device(sbm_engine* e, float** out_device_ptr) {
	if (!e || !out_device_ptr) return SBM_ERR_INVALID_ARG;
	// the block of the buffer finished (or submitted) last
	*out_device_ptr = (float*)((unsigned char*)e->dv_out + (size_t)((e->buffer_index + 1) & 1) * e->out_bytes);
	return SBM_OK;
}
SBM_API void* sbm_stream(sbm_engine* e) { return e ? e->stream : nullptr; }

SBM_API int sbm_poll_finished(sbm_engine* e, int* ids, int cap, int* out_n) {
	if (!e || !ids || cap < 0 || !out_n) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
	const int n = std::min<int>(cap, (int)e->finished_queue.size());
	for (int i = 0; i < n; ++i) ids[i] = e->finished_queue[(size_t)i];
	e->finished_queue.erase(e->finished_queue.begin(), e->finished_queue.begin() + n);
	*out_n = n;
	return SBM_OK;
}

SBM_API int sbm_source_query(sbm_engine* e, int id, sbm_source_state* out) {
	if (!e || !out) return SBM_ERR_INVALID_ARG;
	memset(out, 0, sizeof(*out));
	unsigned long long rd = 0, written = 0; long long eof = 0; float dg = 0, occ = 0; double pos = 0; uint32_t live = 0;
	bool ring = false;
	{
		std::lock_guard<std::mutex> lk(e->mu);
		dev::use(e->cfg.device);
		const int slot = slot_for(e, id);
		if (slot < 0) return SBM_ERR_UNKNOWN_ID;
		const SrcStatic& st = e->h_stat[(uint32_t)slot];
		int rc = dev::d2h(&rd, e->A.read_i + slot, 8, e->stream) | dev::d2h(&eof, e->A.eof + slot, 8, e->stream) |
		         dev::d2h(&dg, e->A.dgain + slot, 4, e->stream) | dev::d2h(&occ, e->A.occ_cur + slot, 4, e->stream) |
		         dev::d2h(&live, e->A.live + slot, 4, e->stream);
		if (st.nz) rc |= dev::d2h(&pos, e->A.mix_pos + st.mix_first, 8, e->stream);
		if (rc) return fail_cuda(e, "sbm_source_query");
		ring = ((st.flags >> kFlagFeedShift) & 3) == kFeedRing;
		written = e->rings[(uint32_t)slot].written;
	}
	if (dev::sync(e->stream)) return fail_cuda(e, "sbm_source_query");     // waits for the buffers in flight, without the engine lock
	out->cur_read_i = ring ? 0 : rd;
	out->ring_size = ring ? written - rd : 0;
	out->eof_marker = eof; out->mix_pos0 = pos; out->direct_gain = dg; out->occlusion = occ; out->alive = (int)live;
	return SBM_OK;
}

SBM_API int sbm_get_stats(sbm_engine* e, sbm_stats* out) {
	if (!e || !out) return SBM_ERR_INVALID_ARG;
	std::lock_guard<std::mutex> lk(e->mu);
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
