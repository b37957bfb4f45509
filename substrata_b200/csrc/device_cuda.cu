// device_cuda.cu -- CUDA implementation of the device seam (device_api.h): memory, copies, events, and "K4", the
// per-buffer finish kernel (SH-domain HRIR decode of the 16-channel bus + stereo mix + interleave + bookkeeping tail).
// There is no host fallback anywhere in this library: every function reports failure when the CUDA call fails.
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>

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
// K4: one CTA.  Thread i handles frame i (frames <= 1024).
//   1. thread 0 decides what happens this buffer (plan_finish) from the -- possibly cross-rank summed -- bus flags;
//   2. decode: out_L/R[i] = sum_c (+-) sum_t h_c[t] * bus_c[i - t], history from the previous buffers (k2_core.h);
//      16 x taps MACs per frame with the taps and the bus staged through shared memory by channel;
//   3. FIR history <- last taps-1 samples of {history, bus};
//   4. out[2i], out[2i+1] = stereo bus (panned + non-spatial sources) + decoder output; tail = {valid, finished, counters}
//      followed by the read counters of ring-fed sources (for CircularBuffer::size(), sbm_source_ring_size).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k4_finish_kernel(FinishArgs a) {
	extern __shared__ float smem[];   // [taps-1 + frames] input of the current channel, [taps] its HRIR
	__shared__ FinishPlan plan;
	const uint32_t F = a.frames, i = threadIdx.x, T = blockDim.x;
	const uint32_t taps = a.hrir ? a.taps : 0;
	const int H = (int)taps - 1;
	if (i == 0) plan = plan_finish(a.bus + (size_t)kBusRows * F, a.hrir != nullptr, a.flags, a.parity, F);
	__syncthreads();
	float l = 0.0f, r = 0.0f;
	if (plan.decode) {
		float* x = smem;                       // x[H + j] = bus_c[j], x[0..H) = history
		float* h = smem + (H > 0 ? H : 0) + F;
		for (int c = 0; c < kNumSh; ++c) {
			const float* bus_c = a.bus + (size_t)c * F;
			float* hist_c = a.hist + (size_t)c * (size_t)(H > 0 ? H : 1);
			__syncthreads();
			for (uint32_t q = i; q < (uint32_t)(H > 0 ? H : 0); q += T) x[q] = hist_c[q];
			for (uint32_t q = i; q < F; q += T) x[(H > 0 ? H : 0) + q] = bus_c[q];
			for (uint32_t q = i; q < taps; q += T) h[q] = a.hrir[(size_t)c * taps + q];
			__syncthreads();
			if (i < F) {
				double acc = 0.0;
				const float* xi = x + (H > 0 ? H : 0) + i;
				for (uint32_t t = 0; t < taps; ++t) acc = fma((double)h[t], (double)xi[-(int)t], acc);
				const float y = (float)acc;
				l += y;
				r = sh_right_ear_negated(c) ? r - y : r + y;
			}
			// carry the tail: the last H samples of {history, bus}
			for (uint32_t q = i; q < (uint32_t)(H > 0 ? H : 0); q += T) hist_c[q] = x[F + q];
		}
	}
	if (i < F) {
		if (plan.stereo) { l += a.bus[(size_t)16 * F + i]; r += a.bus[(size_t)17 * F + i]; }
		if (!plan.valid) { l = 0.0f; r = 0.0f; }
		a.out[2 * i] = l; a.out[2 * i + 1] = r;
	}
	OutTail* tail = reinterpret_cast<OutTail*>(a.out + 2 * F);
	if (i == 0) {
		tail->valid = plan.valid;
		tail->n_finished = a.flags->n_finished;
		tail->n_rendered = a.flags->n_rendered;
		tail->pad = 0;
		tail->algo_bytes = a.flags->algo_bytes;
		a.flags->n_finished = 0; a.flags->n_rendered = 0; a.flags->algo_bytes = 0;
	}
	unsigned long long* reads = reinterpret_cast<unsigned long long*>(tail + 1);
	for (uint32_t q = i; q < a.n_ring; q += T) reads[q] = a.read_i[a.list_ring[q]];
}

namespace dev {

void note_launch(unsigned n) { g_launches += n; }

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
int h2d(void* dst, const void* src, size_t bytes, void* stream) {
	return bytes ? check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream), "cudaMemcpyAsync(H2D)") : 0;
}
int d2h(void* dst, const void* src, size_t bytes, void* stream) {
	return bytes ? check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "cudaMemcpyAsync(D2H)") : 0;
}
int zero(void* dst, size_t bytes, void* stream) { return check(cudaMemsetAsync(dst, 0, bytes, (cudaStream_t)stream), "cudaMemsetAsync"); }
int sync(void* stream) { return check(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize"); }

int launch_finish(const FinishArgs& a, void* stream) {
	const uint32_t F = a.frames;
	const uint32_t threads = ((F + 31) / 32) * 32;
	const uint32_t taps = a.hrir ? a.taps : 0;
	const size_t smem = sizeof(float) * ((size_t)(taps ? taps - 1 : 0) + F + taps);
	k4_finish_kernel<<<1, threads, smem, (cudaStream_t)stream>>>(a);
	note_launch();
	return check_launch("k4_finish");
}

void* event_create() {
	cudaEvent_t e;
	return cudaEventCreate(&e) == cudaSuccess ? (void*)e : nullptr;
}
void event_destroy(void* ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }
void event_record(void* ev, void* stream) { if (ev) cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream); }
float event_elapsed_ms(void* a, void* b) {
	float ms = 0.0f;
	if (!a || !b || cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b) != cudaSuccess) return -1.0f;
	return ms;
}
unsigned long long launches() { return g_launches.load(); }

}  // namespace dev
}  // namespace sbm
