// Which SM sub-partition (scheduler) does warp w of a CTA land on?  Prints %smid / %warpid for 160-thread CTAs, 3 per SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(160, 3) probe(unsigned* out) {
	extern __shared__ float pad[];
	unsigned smid, warpid;
	asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
	asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
	if ((threadIdx.x & 31) == 0) { out[(blockIdx.x * 5 + threadIdx.x / 32) * 2] = smid; out[(blockIdx.x * 5 + threadIdx.x / 32) * 2 + 1] = warpid; }
	// keep the CTA resident long enough for all CTAs of the wave to coexist
	long long t0 = clock64();
	while (clock64() - t0 < 2000000) { }
	if (pad[threadIdx.x] == 123.f) out[0] = 0;
}
int main() {
	const int grid = 148 * 3;
	unsigned* d; cudaMalloc(&d, grid * 5 * 2 * sizeof(unsigned));
	cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
	probe<<<grid, 160, 60000>>>(d);
	unsigned* h = new unsigned[grid * 10];
	cudaMemcpy(h, d, grid * 10 * sizeof(unsigned), cudaMemcpyDeviceToHost);
	printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
	for (int b = 0; b < grid; ++b) if (h[b * 10] < 3) { printf("cta %3d sm %u warpids:", b, h[b * 10]); for (int w = 0; w < 5; ++w) printf(" %u", h[(b * 5 + w) * 2 + 1]); printf("\n"); }
	return 0;
}
