// Microbenchmark: which lanes of a warp share a shared-memory wavefront for LDS.128 gathers?
// For each partner lane b, lanes {0, b} read 16-byte chunks in the SAME bank-quad column but
// different rows (a guaranteed bank conflict iff they are served by the same wavefront).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_pair(uint32_t partner, int iters, long long *out, uint32_t *sink, int mode) {
	extern __shared__ __align__(128) uint8_t smem[];
	for (int i = threadIdx.x; i < 16384; i += blockDim.x) ((uint32_t *)smem)[i] = i;
	__syncthreads();
	uint32_t lane = threadIdx.x;
	bool active = mode == 0 ? (lane == 0 || lane == partner) : true;
	uint32_t col, row;
	if (mode == 0) { col = 0; row = lane; }
	else if (mode == 1) { col = lane & 7; row = (lane * 37 + 11) & 255; }                    // consecutive-8 groups
	else if (mode == 2) { col = (lane >> 2) & 7; row = (lane * 37 + 11) & 255; }             // stride-4 groups
	else if (mode == 3) { col = (lane & 3) | (((lane >> 4) & 1) << 2); row = (lane * 37 + 11) & 255; }
	else if (mode == 4) { col = (lane & 1) | (((lane >> 3) & 3) << 1); row = (lane * 37 + 11) & 255; }
	else if (mode == 5) { col = lane & 7; row = 5; }                                           // all same row: ideal
	else { col = (lane * 5) & 7; row = (lane * 37 + 11) & 255; }
	uint32_t addr = row * 128 + col * 16;
	uint4 acc = make_uint4(0, 0, 0, 0);
	__syncthreads();
	long long t0 = clock64();
	if (active) {
		for (int i = 0; i < iters; i++) {
			uint4 v = *(const uint4 *)(smem + addr);
			acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
			addr = (addr + (v.x & 0) + 128 * 8) & 32767;   // dependent chain keeps loads ordered, same column
		}
	}
	long long t1 = clock64();
	if (lane == 0) out[0] = t1 - t0;
	sink[threadIdx.x] = acc.x ^ acc.y ^ acc.z ^ acc.w;
}
// throughput version: many warps, independent loads
__global__ void k_tp(int iters, long long *out, uint32_t *sink, int mode) {
	extern __shared__ __align__(128) uint8_t smem[];
	for (int i = threadIdx.x; i < 16384; i += blockDim.x) ((uint32_t *)smem)[i] = i * 2654435761u;
	__syncthreads();
	uint32_t lane = threadIdx.x & 31;
	uint32_t col;
	if (mode == 1) col = lane & 7;
	else if (mode == 2) col = (lane >> 2) & 7;
	else if (mode == 3) col = (lane & 3) | (((lane >> 4) & 1) << 2);
	else if (mode == 4) col = (lane & 1) | (((lane >> 3) & 3) << 1);
	else if (mode == 6) col = (lane >> 1) & 7;
	else if (mode == 7) col = ((lane >> 1) & 3) | (((lane >> 4) & 1) << 2);
	else col = lane & 7;
	uint32_t r = (threadIdx.x * 2654435761u) >> 8;
	uint4 acc = make_uint4(0, 0, 0, 0);
	__syncthreads();
	long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 8; u++) {
			r = r * 1664525u + 1013904223u;
			uint32_t row = mode == 5 ? 5 : ((r >> 10) & 255);
			uint4 v = *(const uint4 *)(smem + row * 128 + col * 16);
			acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
		}
	}
	__syncthreads();
	long long t1 = clock64();
	if (threadIdx.x == 0) out[0] = t1 - t0;
	sink[threadIdx.x] = acc.x ^ acc.y ^ acc.z ^ acc.w;
}
int main() {
	long long *d_out; uint32_t *sink;
	cudaMalloc(&d_out, 8); cudaMalloc(&sink, 4096 * 4);
	cudaFuncSetAttribute(k_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
	cudaFuncSetAttribute(k_tp, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
	long long h;
	int iters = 2000;
	printf("pair test (lane 0 + partner, same column different row): cycles/iter\n");
	for (uint32_t b = 1; b < 32; b++) {
		k_pair<<<1, 32, 65536>>>(b, iters, d_out, sink, 0);
		cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
		printf("  partner %2u: %.2f\n", b, (double)h / iters);
	}
	printf("throughput test, 512 threads, 8 LDS.128 per iter: cycles per warp-LDS (ideal 4)\n");
	for (int mode : {5, 1, 2, 3, 4, 6, 7}) {
		k_tp<<<1, 512, 65536>>>(500, d_out, sink, mode);
		cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
		printf("  mode %d: %.2f cycles per warp-level LDS.128\n", mode, (double)h / (500.0 * 8 * 16));
	}
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
