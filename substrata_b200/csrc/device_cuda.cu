// device_cuda.cu -- CUDA implementation of the device seam (device_api.h): memory, copies, events, and "K4": the
// SH-domain HRIR decode of the 16-channel bus (K4a) and the per-buffer finish kernel (K4b: stereo mix + interleave +
// FIR history + bookkeeping tail).
// There is no host fallback anywhere in this library: every function reports failure when the CUDA call fails.
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <type_traits>

#include "device_internal.h"
#include "k2_core.h"

namespace sbm {

namespace {
std::atomic<unsigned long long> g_launches{0};
std::mutex g_err_mu;
std::string g_err = "no error";
thread_local std::string t_err_copy;
}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// K4a: SH-domain HRIR decode, 32 frames per CTA, one warp per ambisonic channel (16 warps): thread (c, f) computes
//   y_c[i] = sum_t h_c[t] * bus_c[i - t]   (double accumulation; history from the previous buffers, k2_core.h)
// from shared-memory copies of the channel's taps and of the input span [i0 - taps + 1, i0 + 32); the 16 channel results
// of a frame are then summed in ACN order into L and, sign by ACN degree, into R (ambisonic_binaural_decoder.cc:60-76).
// Whether this buffer decodes at all is the read-only half of plan_finish (K4b makes the same decision and records it).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDecFrames = 32;
// DBL: taps and inputs are converted to double ONCE while they are staged (f32 -> f64 conversions run on the 16-lane XU
// pipe: two per tap and thread made this kernel 16 us long); the float staging is the fall-back when 8-byte copies of
// [16][taps] + [16][taps + 31] do not fit in shared memory (taps > ~430).
// ORDER1: the same decode for the reflections' first-order buffer (4 channels, the order-1 SH-HRIRs, its own history and flush
// counter) -- room effects.
template <bool DBL, bool ORDER1>
__global__ void __launch_bounds__(kNumSh * 32) k4a_decode_kernel(FinishArgs a) {
	typedef typename std::conditional<DBL, double, float>::type stage_t;
	extern __shared__ __align__(16) unsigned char smem_raw[];   // [16][taps] HRIRs | [16][taps - 1 + 32] inputs | [16][32] channel results (float)
	constexpr int NCH = ORDER1 ? 4 : kNumSh;
	const uint32_t F = a.frames, taps = ORDER1 ? a.taps1 : a.taps;
	const int H = (int)taps - 1;
	grid_dep_wait();
	const float* bf = a.bus + (size_t)kBusRows * F;
	if (ORDER1) { if (!room_decode1_wanted(a.flags, F)) return; }
	else if (!(bf[0] > 0.0f) && !(a.flags->sh_flush_frames < (int)F)) return;      // nothing to decode, no tail to flush
	const uint32_t c = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	const uint32_t i0 = blockIdx.x * kDecFrames;
	const uint32_t span = (uint32_t)H + kDecFrames;
	stage_t* sm = reinterpret_cast<stage_t*>(smem_raw);
	stage_t* h = sm + (size_t)c * taps;
	stage_t* x = sm + (size_t)NCH * taps + (size_t)c * span;
	float* ys = reinterpret_cast<float*>(sm + (size_t)NCH * taps + (size_t)NCH * span);
	// order 1: the reflections' buffer -- all zeros while the decoder only flushes its tail (the node produced nothing)
	const bool silent_in = ORDER1 && !a.flags->foa_valid;
	const float* bus_c = (ORDER1 ? a.foa : a.bus) + (size_t)c * F;
	const float* hist_c = (ORDER1 ? a.hist1 : a.hist) + (size_t)c * (size_t)(H > 0 ? H : 1);
	const float* hrir = ORDER1 ? a.hrir1 : a.hrir;
	for (uint32_t q = lane; q < taps; q += 32) h[q] = (stage_t)hrir[(size_t)c * taps + q];
	for (uint32_t q = lane; q < span; q += 32) {
		const int j = (int)i0 - H + (int)q;                        // input sample index; negative = history
		float v = 0.0f;
		if (j < (int)F) v = (silent_in && j >= 0) ? 0.0f : decode_input(bus_c, hist_c, H, j);
		x[q] = (stage_t)v;
	}
	__syncwarp();
	// four interleaved partial sums break the taps' dependent DFMA chain; everything is accumulated in double, so the
	// summation order does not show in the float result
	double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
	const stage_t* xi = x + H + lane;
	uint32_t t = 0;
	for (; t + 4 <= taps; t += 4) {
		acc0 = fma((double)h[t], (double)xi[-(int)t], acc0);
		acc1 = fma((double)h[t + 1], (double)xi[-(int)t - 1], acc1);
		acc2 = fma((double)h[t + 2], (double)xi[-(int)t - 2], acc2);
		acc3 = fma((double)h[t + 3], (double)xi[-(int)t - 3], acc3);
	}
	for (; t < taps; ++t) acc0 = fma((double)h[t], (double)xi[-(int)t], acc0);
	ys[c * kDecFrames + lane] = (float)((acc0 + acc1) + (acc2 + acc3));
	__syncthreads();
	if (c == 0 && i0 + lane < F) {
		float l = 0.0f, r = 0.0f;
#pragma unroll
		for (int cc = 0; cc < NCH; ++cc) {
			const float y = ys[cc * kDecFrames + lane];
			l += y;
			r = sh_right_ear_negated(cc) ? r - y : r + y;
		}
		float* dec = ORDER1 ? a.dec1 : a.dec;
		dec[i0 + lane] = l; dec[F + i0 + lane] = r;
	}
}

// ---------------------------------------------------------------------------------------------------------------
// Room kernel: Resonance's ReflectionsNode for this buffer (room_core.h).  One CTA, thread i = frame i.  Thread 0 plans
// (gain processors, crossfade, rotation mode, tail bookkeeping), then every frame is independent: insert the (low-passed)
// reflections bus into the delay line, read six delayed copies, ramp, encode to first order, crossfade, rotate.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_room_kernel(RoomArgs a) {
	__shared__ RoomPlan plan;
	__shared__ float lp[kMaxFrames];
	const uint32_t F = a.frames, i = threadIdx.x;
	grid_dep_wait();
	const float* in = a.bus + (size_t)kRowRefl * F;
	if (i == 0) {
		const RoomParams p = *a.params;
		const bool have_input = p.enabled && a.bus[(size_t)kBusRows * F + 3] > 0.0f;
		room_plan(p, a.state, have_input, F, &plan);
		if (plan.process && !(plan.lp_coef < kEpsilonFloat)) {
			// MonoPoleFilter: a serial recurrence over the buffer (only with a cut-off frequency set; Substrata sets none)
			float prev = plan.lp_prev;
			for (uint32_t n = 0; n < F; ++n) { prev = lpf_step(plan.lp_coef, prev, plan.silent_input ? 0.0f : in[n]); lp[n] = prev; }
			a.state->lp_prev = prev;
		} else if (plan.process) {
			a.state->lp_prev = plan.silent_input ? 0.0f : in[F - 1];      // bypassed filter remembers its last input (mono_pole_filter.cc:40-44)
		}
		a.flags->foa_valid = plan.process ? 1 : 0;
	}
	__syncthreads();
	if (!plan.process) return;
	if (i < F) {
		const float v = !(plan.lp_coef < kEpsilonFloat) ? lp[i] : (plan.silent_input ? 0.0f : in[i]);
		a.delay_line[room_insert_index(plan, F, i)] = v;
	}
	__syncthreads();      // (one CTA: the inserted frames are visible to every thread's reads below)
	if (i < F) {
		float w, y, z, x;
		room_frame(plan, a.delay_line, F, i, &w, &y, &z, &x);
		a.foa[i] = w; a.foa[(size_t)F + i] = y; a.foa[(size_t)2 * F + i] = z; a.foa[(size_t)3 * F + i] = x;
	}
}

// ---------------------------------------------------------------------------------------------------------------
// K4b: one CTA.  Thread i handles frame i (frames <= 1024).
//   1. thread 0 decides what happens this buffer (plan_finish) from the -- possibly cross-rank summed -- bus flags;
//   2. FIR history <- last taps-1 samples of {history, bus} (after K4a has read the old history);
//   3. out[2i], out[2i+1] = stereo bus (panned + non-spatial sources) + decoder output; tail = {valid, finished, counters}
//      followed by the read counters of ring-fed sources (for CircularBuffer::size(), sbm_source_ring_size).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k4b_finish_kernel(FinishArgs a) {
	extern __shared__ float smem[];   // [taps - 1] new history of the current channel
	__shared__ FinishPlan plan;
	const uint32_t F = a.frames, i = threadIdx.x, T = blockDim.x;
	const uint32_t taps = a.hrir ? a.taps : 0;
	const int H = (int)taps - 1;
	grid_dep_wait();
	const int foa_valid = a.hrir1 ? a.flags->foa_valid : 0;      // (read by every thread before thread 0's plan_finish clears it: barrier)
	__syncthreads();
	if (i == 0) plan = plan_finish(a.bus + (size_t)kBusRows * F, a.hrir != nullptr, a.hrir1 != nullptr, a.room_enabled != 0, a.flags, a.parity, F);
	__syncthreads();
	float l = 0.0f, r = 0.0f;
	if (plan.decode1) {
		// the reflections' order-1 decode (K4a<.., true>) + its FIR history: last taps1 - 1 samples of {history, buffer or silence}
		if (i < F) { l = a.dec1[i]; r = a.dec1[F + i]; }
		const int H1 = (int)a.taps1 - 1;
		if (H1 > 0) {
			const uint32_t n1 = 4u * (uint32_t)H1;      // <= 4 * 1023 floats: through shared memory (it may overlap the old history when frames < taps)
			for (uint32_t q = i; q < n1; q += T) {
				const uint32_t c = q / (uint32_t)H1, k = q - c * (uint32_t)H1;
				const int j = (int)F - H1 + (int)k;
				smem[q] = (j >= 0 && !foa_valid) ? 0.0f : decode_input(a.foa + (size_t)c * F, a.hist1 + (size_t)c * (size_t)H1, H1, j);
			}
			__syncthreads();
			for (uint32_t q = i; q < n1; q += T) a.hist1[q] = smem[q];
			__syncthreads();
		}
	}
	if (plan.decode) {
		if (i < F) { l += a.dec[i]; r += a.dec[F + i]; }
		if (H > 0 && a.hist_smem_all) {
			// all 16 channels at once: the new history goes through shared memory (it may overlap the old one when frames < taps)
			const uint32_t n = (uint32_t)kNumSh * (uint32_t)H;
			for (uint32_t q = i; q < n; q += T) {
				const uint32_t c = q / (uint32_t)H, k = q - c * (uint32_t)H;
				smem[q] = decode_input(a.bus + (size_t)c * F, a.hist + (size_t)c * (size_t)H, H, (int)F - H + (int)k);
			}
			__syncthreads();
			for (uint32_t q = i; q < n; q += T) a.hist[q] = smem[q];
		} else if (H > 0) {
			for (int c = 0; c < kNumSh; ++c) {
				const float* bus_c = a.bus + (size_t)c * F;
				float* hist_c = a.hist + (size_t)c * (size_t)H;
				__syncthreads();
				for (uint32_t q = i; q < (uint32_t)H; q += T) smem[q] = decode_input(bus_c, hist_c, H, (int)F - H + (int)q);
				__syncthreads();
				for (uint32_t q = i; q < (uint32_t)H; q += T) hist_c[q] = smem[q];
			}
		}
	}
	if (i < F) {
		if (plan.stereo) { l += a.bus[(size_t)16 * F + i]; r += a.bus[(size_t)17 * F + i]; }
		if (!plan.valid) { l = 0.0f; r = 0.0f; }
		a.out[2 * i] = l; a.out[2 * i + 1] = r;
	}
	OutTail* tail = reinterpret_cast<OutTail*>(a.out + 2 * F);
	// K1's accounting slots: one lane each (kStatParts == 32), summed by shuffles
	unsigned long long part_bytes = 0; unsigned int part_rendered = 0;
	if (i < 32) {
		part_bytes = a.flags->part_bytes[i][0]; part_rendered = a.flags->part_rendered[i][0];
		a.flags->part_bytes[i][0] = 0; a.flags->part_rendered[i][0] = 0;
		for (int o = 16; o > 0; o >>= 1) { part_bytes += __shfl_xor_sync(0xffffffffu, part_bytes, o); part_rendered += __shfl_xor_sync(0xffffffffu, part_rendered, o); }
	}
	const int n_fin = a.flags->n_finished;     // (read by every thread before thread 0 clears it below: barrier first)
	__syncthreads();
	if (i == 0) {
		tail->valid = plan.valid;
		tail->n_finished = n_fin;
		tail->n_rendered = a.flags->n_rendered + part_rendered;
		tail->status = a.flags->pipe_timeout ? kTailPipeTimeout : (a.flags->peer_timeout ? kTailPeerTimeout : kTailOk);
		tail->algo_bytes = a.flags->algo_bytes + part_bytes;
		a.flags->n_finished = 0; a.flags->n_rendered = 0; a.flags->algo_bytes = 0; a.flags->peer_timeout = 0; a.flags->pipe_timeout = 0;
	}
	unsigned long long* reads = reinterpret_cast<unsigned long long*>(tail + 1);
	for (uint32_t q = i; q < a.n_ring; q += T) reads[q] = a.read_i[a.list_ring[q]];
	// the ids of sources the feeder stopped this buffer ride along (engine.cc fetches the list separately only when it is longer)
	int* fin = reinterpret_cast<int*>(reads + a.n_ring);
	for (int q = (int)i; q < n_fin && q < kInlineFinished; q += (int)T) fin[q] = a.finished[q];
}

// Batched ring push (the streaming feed, AE.cpp:540-545 / ClientUDPHandlerThread.cpp:239-242 for many sources at once):
// the packets of one call arrive as ONE staged payload (single H2D from pinned memory); one CTA per packet copies it
// into its source's ring with wrap-around.  Coalesced 4-byte accesses; packets are a few hundred samples.
__global__ void __launch_bounds__(128) k_ring_scatter_kernel(const RingPush* __restrict__ pk, const float* __restrict__ stage) {
	const RingPush p = pk[blockIdx.x];
	for (uint32_t i = threadIdx.x; i < p.n; i += blockDim.x) {
		uint32_t at = p.at + i;
		if (at >= p.capacity) at -= p.capacity;
		if (at >= p.capacity) at %= p.capacity;
		p.ring[at] = stage[p.src_off + i];
	}
}

// Applies one upload block (engine.cc): blockIdx.y = segment; the segment's words are copied to its device array, element j
// to dst[j] (contiguous) or dst[idx[j]] (scattered setters).  Word-granular so that one kernel serves 4- to 32-byte elements.
__global__ void __launch_bounds__(256) k_apply_uploads_kernel(const UpSeg* __restrict__ segs, const uint32_t* __restrict__ block) {
	const UpSeg sg = segs[blockIdx.y];
	uint32_t* __restrict__ dst = reinterpret_cast<uint32_t*>(sg.dst);
	const uint32_t* __restrict__ val = block + sg.src_word;
	const uint32_t ew = sg.elem_words;
	const size_t total = (size_t)sg.count * ew;
	for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
		const uint32_t j = (uint32_t)(w / ew), k = (uint32_t)(w - (size_t)j * ew);
		const uint32_t at = sg.idx_word == 0xffffffffu ? j : block[sg.idx_word + j];
		dst[(size_t)at * ew + k] = val[w];
	}
}

namespace dev {

void note_launch(unsigned n) { g_launches += n; }

int launch_apply_uploads(const UpSeg* segs, uint32_t n_segs, const uint32_t* block, uint32_t max_words, void* stream) {
	if (!n_segs) return 0;
	uint32_t gx = (max_words + 1023) / 1024;      // <= 4 words per thread
	if (gx < 1) gx = 1;
	if (gx > 592) gx = 592;
	k_apply_uploads_kernel<<<dim3(gx, n_segs), 256, 0, (cudaStream_t)stream>>>(segs, block);
	note_launch();
	return check_launch("k_apply_uploads");
}

int check(cudaError_t e, const char* what) {
	if (e == cudaSuccess) return 0;
	std::lock_guard<std::mutex> lk(g_err_mu);
	g_err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
	return -1;
}
int check_launch(const char* what) { return check(cudaGetLastError(), what); }

const char* last_error() {
	std::lock_guard<std::mutex> lk(g_err_mu);
	t_err_copy = g_err;
	return t_err_copy.c_str();
}

int use(int device) { return check(cudaSetDevice(device), "cudaSetDevice"); }

int init(int device, void** stream_out) {
	int count = 0;
	if (check(cudaGetDeviceCount(&count), "cudaGetDeviceCount")) return -1;
	if (device < 0 || device >= count) {
		std::lock_guard<std::mutex> lk(g_err_mu);
		g_err = "no such CUDA device";
		return -1;
	}
	if (use(device)) return -1;
	cudaDeviceProp prop;
	if (check(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return -1;
	if (prop.major != 10) {
		std::lock_guard<std::mutex> lk(g_err_mu);
		g_err = std::string("libsubstrata_b200 is built for sm_100a only; device is ") + prop.name;
		return -1;
	}
	cudaStream_t s;
	if (check(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate")) return -1;
	*stream_out = (void*)s;
	return 0;
}

void* stream_create() {
	cudaStream_t s;
	return check(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate") ? nullptr : (void*)s;
}

void shutdown(void* stream) { if (stream) cudaStreamDestroy((cudaStream_t)stream); }

void* alloc(size_t bytes) {
	void* p = nullptr;
	if (check(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc")) return nullptr;
	// cudaMemset on device memory is asynchronous to the host and runs on the legacy default stream, which the engine's
	// non-blocking streams do not wait for: finish it here so that a following upload cannot be overwritten by zeros.
	if (check(cudaMemset(p, 0, bytes ? bytes : 1), "cudaMemset") || check(cudaStreamSynchronize(0), "cudaMemset(sync)")) { cudaFree(p); return nullptr; }
	return p;
}
void release(void* p) { if (p) cudaFree(p); }
void* alloc_pinned(size_t bytes) {
	void* p = nullptr;
	if (check(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault), "cudaHostAlloc")) return nullptr;
	memset(p, 0, bytes);
	return p;
}
void release_pinned(void* p) { if (p) cudaFreeHost(p); }
bool is_pinned(const void* p) {
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return at.type == cudaMemoryTypeHost;
}
int h2d(void* dst, const void* src, size_t bytes, void* stream) {
	return bytes ? check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream), "cudaMemcpyAsync(H2D)") : 0;
}
int d2h(void* dst, const void* src, size_t bytes, void* stream) {
	return bytes ? check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "cudaMemcpyAsync(D2H)") : 0;
}
int zero(void* dst, size_t bytes, void* stream) { return check(cudaMemsetAsync(dst, 0, bytes, (cudaStream_t)stream), "cudaMemsetAsync"); }
int sync(void* stream) { return check(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize"); }

int peer_export(void* device_ptr, void* handle_out64) {
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
	cudaIpcMemHandle_t h;
	if (check(cudaIpcGetMemHandle(&h, device_ptr), "cudaIpcGetMemHandle")) return -1;
	memcpy(handle_out64, &h, sizeof(h));
	return 0;
}
void* peer_open(const void* handle64) {
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, sizeof(h));
	void* p = nullptr;
	if (check(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return nullptr;
	return p;
}
void peer_close(void* mapped) { if (mapped) cudaIpcCloseMemHandle(mapped); }

int launch_ring_scatter(const RingPush* packets, uint32_t n, const float* stage, void* stream) {
	if (!n) return 0;
	k_ring_scatter_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(packets, stage);
	note_launch();
	return check_launch("k_ring_scatter");
}

int launch_room(const RoomArgs& a, void* stream) {
	const uint32_t threads = ((a.frames + 31) / 32) * 32;
	launch_pdl(k_room_kernel, dim3(1), dim3(threads), 0, (cudaStream_t)stream, a);
	note_launch();
	return check_launch("k_room");
}

int launch_finish(const FinishArgs& a, void* stream) {
	const uint32_t F = a.frames;
	cudaStream_t st = (cudaStream_t)stream;
	if (a.hrir) {
		const size_t elems = (size_t)kNumSh * a.taps + (size_t)kNumSh * (a.taps - 1 + kDecFrames);
		const bool dbl = sizeof(double) * elems + sizeof(float) * kNumSh * kDecFrames <= 200 * 1024;
		const size_t smem = (dbl ? sizeof(double) : sizeof(float)) * elems + sizeof(float) * (size_t)kNumSh * kDecFrames;
		static size_t configured[2] = {48 * 1024, 48 * 1024};
		if (smem > configured[dbl]) {
			if (check(dbl ? cudaFuncSetAttribute(k4a_decode_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
			              : cudaFuncSetAttribute(k4a_decode_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(k4a)")) return -1;
			configured[dbl] = smem;
		}
		if (dbl) launch_pdl(k4a_decode_kernel<true, false>, dim3((F + kDecFrames - 1) / kDecFrames), dim3(kNumSh * 32), smem, st, a);
		else launch_pdl(k4a_decode_kernel<false, false>, dim3((F + kDecFrames - 1) / kDecFrames), dim3(kNumSh * 32), smem, st, a);
		note_launch();
		if (check_launch("k4a_decode")) return -1;
	}
	if (a.hrir1) {
		// the reflections' order-1 decode: 4 channels, float staging (4 x (2 taps + 31) floats <= 33 KB)
		const size_t elems = (size_t)4 * a.taps1 + (size_t)4 * (a.taps1 - 1 + kDecFrames);
		const size_t smem = sizeof(float) * elems + sizeof(float) * (size_t)4 * kDecFrames;
		launch_pdl(k4a_decode_kernel<false, true>, dim3((F + kDecFrames - 1) / kDecFrames), dim3(4 * 32), smem, st, a);
		note_launch();
		if (check_launch("k4a_decode(order 1)")) return -1;
	}
	const uint32_t threads = ((F + 31) / 32) * 32;
	const size_t h1 = a.hrir && a.taps > 1 ? a.taps - 1 : 1;
	FinishArgs b = a;
	b.hist_smem_all = sizeof(float) * h1 * kNumSh <= 48 * 1024 ? 1 : 0;      // 16 channels' history in one pass when it fits the default limit
	size_t smem = sizeof(float) * h1 * (b.hist_smem_all ? (size_t)kNumSh : 1);
	if (a.hrir1 && a.taps1 > 1) smem = smem > sizeof(float) * 4 * (a.taps1 - 1) ? smem : sizeof(float) * 4 * (a.taps1 - 1);
	launch_pdl(k4b_finish_kernel, dim3(1), dim3(threads), smem, st, b);
	note_launch();
	return check_launch("k4b_finish");
}

void* event_create() {
	cudaEvent_t e;
	return cudaEventCreate(&e) == cudaSuccess ? (void*)e : nullptr;
}
void event_destroy(void* ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }
void event_record(void* ev, void* stream) { if (ev) cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream); }
int stream_wait_event(void* stream, void* ev) { return ev ? check(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)ev, 0), "cudaStreamWaitEvent") : 0; }
int event_sync(void* ev) { return ev ? check(cudaEventSynchronize((cudaEvent_t)ev), "cudaEventSynchronize") : 0; }
float event_elapsed_ms(void* a, void* b) {
	float ms = 0.0f;
	if (!a || !b || cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b) != cudaSuccess) return -1.0f;
	return ms;
}
unsigned long long launches() { return g_launches.load(); }

}  // namespace dev
}  // namespace sbm
