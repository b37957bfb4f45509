// device_api.h -- the narrow seam between the engine's host logic (engine.cc) and the device.
//
// The product implements it with CUDA (device_cuda.cu + k*.cu).  tests/native/device_emu.cc implements the same
// functions on the host for the no-GPU CPU test suite (host emulation of the kernels' semantics; test
// infrastructure that is never part of libsubstrata_b200.so).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "sbm_core.h"
#include "room_core.h"

namespace sbm {

struct SlotPatch {              // state (re)initialisation of one slot, applied by k_patch before K1
	uint32_t slot;
	uint32_t what;              // kPatch*
	unsigned long long read_i;
	long long eof;
	int32_t rate;               // resampler src rate (init)
	int32_t pad;
};
constexpr uint32_t kPatchReset = 1;      // fresh source: all DSP state cleared, live = 1
constexpr uint32_t kPatchReadI = 2;      // cur_read_i / ring read counter
constexpr uint32_t kPatchEof = 4;
constexpr uint32_t kPatchResampler = 8;  // AudioResampler::init (AR.cpp:24-36)
constexpr uint32_t kPatchKill = 16;      // live = 0

struct MixPatch { uint32_t index; uint32_t pad; double pos; };

struct RingPush {               // one packet of a batched ring push: stage[src_off .. src_off + n) -> ring[(at + i) % capacity]
	float* ring;
	uint32_t capacity;
	uint32_t at;                // write index of the packet's first sample
	uint32_t src_off;           // offset of the packet in the staged payload, in floats
	uint32_t n;
};

// One segment of the per-buffer upload block (engine.cc: upload_block_locked): `count` elements of `elem_words` 4-byte words,
// values at word `src_word` of the block, copied to dst[0 .. count) (idx_word == 0xffffffff) or scattered to dst[idx[j]].
struct UpSeg {
	void* dst;
	uint32_t elem_words, count;
	uint32_t src_word, idx_word;
	uint32_t pad, pad2;
};
static_assert(sizeof(UpSeg) == 32, "UpSeg is 32 bytes");
constexpr int kMaxUpSegs = 16;
constexpr int kInlineFinished = 1024;   // finished-source ids that ride in the result block (more: fetched separately)

constexpr int kMaxPeers = 8;
// Cross-GPU exchange of the bus over NVLink peer memory (world == 0: single GPU, nothing is exchanged).
// Every rank owns one gather area [2 buffer parities][world][bus_floats] plus two arrival counters; K3 of rank r stores
// its reduced bus straight into slot r of EVERY rank's area (peer stores) and then bumps every rank's counter; the gather
// kernel in front of K4 waits for `world` arrivals and sums the slots in rank order (identical result on all ranks).
struct PeerArgs {
	uint32_t rank, world;
	float* gather[kMaxPeers];           // gather[p]: rank p's area (own memory for p == rank, peer-mapped otherwise)
	unsigned int* arrive[kMaxPeers];    // arrive[p]: rank p's two monotonic arrival counters
	unsigned int* done;                 // local: K3 CTAs finished (last-CTA election)
	uint32_t parity;                    // which half of the gather areas / which arrival counter this buffer uses (exchange count & 1)
	uint32_t pad;
	unsigned long long timeout_ns;      // gather kernel: give up waiting for the peers after this long (the buffer then holds the local bus only)
};

struct MixArgs {
	const MixRec* mrec;         // [n_pan + n_sh] render-list order: stereo-bus sources first, then ambisonic sources (CUDA build)
	const WorkRec* rec;         // indexed like mrec: the full records (the CUDA kernels read them on their careful paths only)
	const ShGains* shg;         // indexed like rec
	const ShRamps* shr;         // indexed like rec
	const ReflSend* rfl;        // indexed like rec; null = room effects off (the kernels then carry no reflections row)
	const MixZ* mixz;
	const double* mix_pos0;
	Resamp* rs;                 // K2 stores prev_samples of resampled sources
	float* lpf_prev;            // K2 stores the occlusion filter memory
	uint32_t n_pan;             // entries rendered by the 2-channel kernel (panned + non-spatial sources)
	uint32_t n_sh;              // entries rendered by the 16-channel kernel
	uint32_t frames;
	int32_t rate;
	float* partials;            // scratch of per-CTA partial buses: [max_ctas(pan)][2][frames] then [max_ctas(sh)][16][frames]
	size_t partials_floats;
	float* bus;                 // [18][frames] + 4 flag floats: this rank's bus
	const DevFlags* flags;
	uint32_t parity;
	void* ev_after_k2;          // optional event recorded between K2 and K3 (profiling)
	PeerArgs peer;
	size_t bus_floats;
	uint32_t rot_div;           // set by the launcher: SM count (CTA index / rot_div = which resident CTA of its SM a CTA is, first wave)
	const uint32_t* ready;      // non-null: K1 is still running beside the mix kernel; records of tile t are valid once ready[t] == epoch
	uint32_t epoch;
};

struct RoomArgs {               // the room kernel (room_core.h): reflections bus -> first-order buffer
	const float* bus;           // (possibly cross-rank summed) [19][frames] + 4 flags: row kRowRefl is the mono reflections send
	const RoomParams* params;   // device copy of this buffer's RoomParams
	RoomState* state;
	float* delay_line;          // [delay_size] mono
	float* foa;                 // [4][frames] out (ACN: W Y Z X)
	DevFlags* flags;
	uint32_t frames;
};

struct FinishArgs {
	const float* bus;           // (possibly all-reduced) [19][frames] + 4 flags
	const float* hrir;          // [16][taps] or null
	uint32_t taps;
	float* hist;                // [16][taps-1] FIR history (the decoder's overlap state), updated in place
	float* dec;                 // [2][frames] scratch: decoder output L / R
	DevFlags* flags;
	uint32_t parity;
	uint32_t frames;
	float* out;                 // [2*frames] interleaved, followed by OutTail
	const uint32_t* list_ring;  // ring-fed slots whose read counters are reported back
	uint32_t n_ring;
	const unsigned long long* read_i;
	const int* finished;        // this buffer's finished-id list (device); the first kInlineFinished ids are copied behind the ring counters
	int hist_smem_all;          // set by launch_finish: K4b updates the FIR history of all 16 channels in one pass
	// room effects: the reflections' first-order buffer and its order-1 SH-HRIR decoder (null hrir1: no room effects configured)
	const float* foa;           // [4][frames]
	const float* hrir1;         // [4][taps1]
	uint32_t taps1;
	float* hist1;               // [4][taps1-1]
	float* dec1;                // [2][frames]
	uint32_t room_enabled;
};

struct OutTail {                // follows the 2*frames output floats in the D2H staging buffer
	int valid;
	int n_finished;
	unsigned int n_rendered;
	int status;                 // kTail*
	unsigned long long algo_bytes;
};
constexpr int kTailOk = 0, kTailPeerTimeout = 1, kTailPipeTimeout = 2;

struct PolyphaseArgs {          // K6 (k6_polyphase.cu): Resonance's polyphase resampler, one launch per Process call
	const float* in;            // [channels][ext]: per channel the carried state (hist = cpp - 1 samples) followed by the input
	const float* coeffs;        // transposed_filter_coeffs_: [up][cpp], each phase flipped
	float* out;                 // [channels][n_out]
	uint32_t channels, ext, hist, n_out;
	uint32_t up, down, time_modulo, last, cpp;
};

namespace dev {
int init(int device, void** stream_out);
void* stream_create();                           // another non-blocking stream on the current device
int use(int device);                             // make `device` current on the calling thread
void shutdown(void* stream);
const char* last_error();
void* alloc(size_t bytes);                       // zero-initialised device memory; nullptr on failure
void release(void* p);
void* alloc_pinned(size_t bytes);
void release_pinned(void* p);
bool is_pinned(const void* p);                   // page-locked host memory known to the CUDA driver (a copy from it needs no staging)
int h2d(void* dst, const void* src, size_t bytes, void* stream);
int d2h(void* dst, const void* src, size_t bytes, void* stream);
int zero(void* dst, size_t bytes, void* stream);
int sync(void* stream);
int num_mix_ctas(uint32_t frames, int sh);       // largest grid K2 will use (sizes the partials buffer)
size_t partials_floats(uint32_t frames);         // floats needed for MixArgs::partials
// kernels
int launch_apply_uploads(const UpSeg* segs, uint32_t n_segs, const uint32_t* block, uint32_t max_words, void* stream);   // copies / scatters every segment of an upload block
int launch_patch(const SlotArrays& A, const SlotPatch* patches, uint32_t n, const MixPatch* mp, uint32_t nmp, void* stream);
int launch_params(const Globals& g, const SlotArrays& A, DevFlags* flags, const float* table, const uint32_t* lists, uint32_t n_slots, void* stream);   // lists: [n_pan + n_sh] slots in render order
int launch_ring_scatter(const RingPush* packets, uint32_t n, const float* stage, void* stream);   // batched ring push: one CTA per packet
int launch_mix(const MixArgs& a, void* stream);  // K2 (both lists) + K3 (partials -> bus)
int launch_polyphase(const PolyphaseArgs& a, void* stream);
int launch_room(const RoomArgs& a, void* stream);   // ReflectionsNode for this buffer (room effects)
int launch_finish(const FinishArgs& a, void* stream);
// peer exchange (multi-GPU): IPC plumbing + the gather kernel
int peer_export(void* device_ptr, void* handle_out64);             // cudaIpcGetMemHandle
void* peer_open(const void* handle64);                             // cudaIpcOpenMemHandle (other process, any GPU of the node)
void peer_close(void* mapped);
int launch_peer_gather(const PeerArgs& pa, const float* own_bus, float* bus_sum, size_t bus_floats, unsigned int target, DevFlags* flags, void* stream);
// timing
void* event_create();
void event_destroy(void* ev);
void event_record(void* ev, void* stream);
int stream_wait_event(void* stream, void* ev);   // work queued on `stream` after this call waits (on the device) for the event
int event_sync(void* ev);                        // blocks the host until the event has happened
float event_elapsed_ms(void* a, void* b);
unsigned long long launches();                   // kernels launched so far by this library
}  // namespace dev
}  // namespace sbm
