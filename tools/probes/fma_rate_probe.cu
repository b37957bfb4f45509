// What bounds the ambisonic mix kernel?  Measures, per SM and clock, the rate of (a) scalar FFMA, (b) packed FFMA2 with the operand
// shapes the kernel uses (gain pair x broadcast sample + accumulator pair), (c) the kernel's inner block -- four broadcast LDS.128 of
// gains feeding 32 FFMA2 -- for 4 .. 32 resident warps per SM.  One CTA per SM, `warps` warps, every thread 32 independent
// accumulator pairs (the kernel's 64 accumulators).   nvcc -O3 -gencode arch=compute_100a,code=sm_100a fma_rate_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& acc, const float2 g, const float2 x) {
	unsigned long long a = *reinterpret_cast<unsigned long long*>(&acc);
	asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(*reinterpret_cast<const unsigned long long*>(&g)), "l"(*reinterpret_cast<const unsigned long long*>(&x)));
	acc = *reinterpret_cast<float2*>(&a);
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(float* out, const float* in, int iters, long long* cycles) {
	__shared__ __align__(16) float gains[64];
	if (threadIdx.x < 64) gains[threadIdx.x] = in[threadIdx.x];
	__syncthreads();
	float2 acc[32];
#pragma unroll
	for (int i = 0; i < 32; ++i) acc[i] = make_float2(0.0f, (float)i);
	float y0 = in[64 + (threadIdx.x & 7)], y1 = y0 + 1.0f, y2 = y0 + 2.0f, y3 = y0 + 3.0f;
	const long long t0 = clock64();
	for (int it = 0; it < iters; ++it) {
		if (MODE == 0) {          // 64 scalar FFMA
#pragma unroll
			for (int i = 0; i < 32; ++i) {
				asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].x) : "f"(y0), "f"(y1));
				asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].y) : "f"(y2), "f"(y3));
			}
		} else if (MODE == 1) {   // 32 FFMA2, register operands
			const float2 g0 = make_float2(y0, y1), g1 = make_float2(y2, y3);
#pragma unroll
			for (int i = 0; i < 32; i += 2) { ffma2(acc[i], g0, make_float2(y1, y1)); ffma2(acc[i + 1], g1, make_float2(y3, y3)); }
		} else if (MODE == 3) {   // 32 FFMA2, both multiplicands genuine pairs (no broadcast operand)
			const float2 g0 = make_float2(y0, y1), g1 = make_float2(y2, y3), x0 = make_float2(y1, y2), x1 = make_float2(y3, y0);
#pragma unroll
			for (int i = 0; i < 32; i += 2) { ffma2(acc[i], g0, x0); ffma2(acc[i + 1], g1, x1); }
		} else if (MODE == 4) {   // the kernel's block with the sample held as a genuine register pair
			const float4* g4 = reinterpret_cast<const float4*>(gains) + (it & 3) * 4;
			const float2 p0 = make_float2(y0, y1), p1 = make_float2(y1, y2), p2 = make_float2(y2, y3), p3 = make_float2(y3, y0);
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const float4 g = g4[q];
				ffma2(acc[8 * q + 0], make_float2(g.x, g.y), p0); ffma2(acc[8 * q + 1], make_float2(g.z, g.w), p0);
				ffma2(acc[8 * q + 2], make_float2(g.x, g.y), p1); ffma2(acc[8 * q + 3], make_float2(g.z, g.w), p1);
				ffma2(acc[8 * q + 4], make_float2(g.x, g.y), p2); ffma2(acc[8 * q + 5], make_float2(g.z, g.w), p2);
				ffma2(acc[8 * q + 6], make_float2(g.x, g.y), p3); ffma2(acc[8 * q + 7], make_float2(g.z, g.w), p3);
			}
		} else if (MODE == 2) {   // the kernel's block: 4 broadcast LDS.128 + 32 FFMA2
			const float4* g4 = reinterpret_cast<const float4*>(gains) + (it & 3) * 4;
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const float4 g = g4[q];
				ffma2(acc[8 * q + 0], make_float2(g.x, g.y), make_float2(y0, y0)); ffma2(acc[8 * q + 1], make_float2(g.z, g.w), make_float2(y0, y0));
				ffma2(acc[8 * q + 2], make_float2(g.x, g.y), make_float2(y1, y1)); ffma2(acc[8 * q + 3], make_float2(g.z, g.w), make_float2(y1, y1));
				ffma2(acc[8 * q + 4], make_float2(g.x, g.y), make_float2(y2, y2)); ffma2(acc[8 * q + 5], make_float2(g.z, g.w), make_float2(y2, y2));
				ffma2(acc[8 * q + 6], make_float2(g.x, g.y), make_float2(y3, y3)); ffma2(acc[8 * q + 7], make_float2(g.z, g.w), make_float2(y3, y3));
			}
		}
	}
	const long long t1 = clock64();
	float s = 0.0f;
#pragma unroll
	for (int i = 0; i < 32; ++i) s += acc[i].x + acc[i].y;
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
	int sms = 0;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	float *out, *in; long long* cyc;
	cudaMalloc(&out, sms * 512 * sizeof(float)); cudaMalloc(&in, 128 * sizeof(float)); cudaMalloc(&cyc, sms * sizeof(long long));
	float h[128]; for (int i = 0; i < 128; ++i) h[i] = 1e-3f * (float)(i + 1);
	cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
	const int iters = 20000;
	const char* names[5] = {"FFMA (64 scalar per iteration)", "FFMA2, one multiplicand a broadcast scalar (32 / it.)", "4 x LDS.128 + 32 FFMA2 (the mix kernel's block)", "FFMA2, both multiplicands register pairs (32 / it.)", "4 x LDS.128 + 32 FFMA2, sample as a register pair"};
	for (int mode = 0; mode < 5; ++mode)
		for (int warps = 2; warps <= 16; warps *= 2) {
			for (int rep = 0; rep < 2; ++rep) {
				if (mode == 0) probe<0><<<sms, warps * 32>>>(out, in, iters, cyc);
				else if (mode == 1) probe<1><<<sms, warps * 32>>>(out, in, iters, cyc);
				else if (mode == 2) probe<2><<<sms, warps * 32>>>(out, in, iters, cyc);
				else if (mode == 3) probe<3><<<sms, warps * 32>>>(out, in, iters, cyc);
				else probe<4><<<sms, warps * 32>>>(out, in, iters, cyc);
				cudaDeviceSynchronize();
			}
			long long c[1024]; cudaMemcpy(c, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
			double mean = 0; for (int i = 0; i < sms; ++i) mean += (double)c[i]; mean /= sms;
			const double fma = 64.0 * 32.0 * warps * iters;     // FMAs per SM
			printf("%-52s warps/SM %2d  cycles %.0f  FMA/clk/SM %.1f  (warp-instr/clk/SM %.2f)\n", names[mode], warps, mean, fma / mean, fma / mean / (mode == 0 ? 32.0 : 64.0));
		}
	printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
