// mb_overlap.cu -- does a pinned host-to-device copy stream overlap with a long shared-memory-bound kernel on another
// stream?  (The streamed univariate round relies on it.)  Prints kernel alone, copies alone, both, for a kernel that
// (a) only spins on shared memory, (b) also reads global memory like k_uni_b8 does.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mb_overlap tools/mb_overlap.cu
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) k_spin(uint32_t *out, const uint32_t *in, uint32_t iters, uint32_t gl_every) {
	extern __shared__ uint32_t sm[];
	for (uint32_t i = threadIdx.x; i < 50 * 1024; i += 512) sm[i] = i * 2654435761u;
	__syncthreads();
	uint32_t x = threadIdx.x, acc = 0;
	for (uint32_t it = 0; it < iters; it++) {
		x = sm[(x + it) % (50 * 1024)];
		acc ^= x;
		if (gl_every && (it % gl_every) == 0) acc ^= __ldg(in + ((blockIdx.x * 512 + threadIdx.x + it) & 0xffffff));
	}
	if (acc == 0x12345678) out[0] = acc;
}

int main() {
	const size_t CH = 2u << 20, NCH = 153;
	char *h, *d;
	uint32_t *din, *dout;
	cudaMallocHost(&h, CH * NCH);
	cudaMalloc(&d, CH * NCH);
	cudaMalloc(&din, 64u << 20);
	cudaMalloc(&dout, 4096);
	cudaStream_t sa, sb;
	cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking);
	cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking);
	cudaFuncSetAttribute(k_spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	cudaEvent_t e0, e1, f0, f1;
	cudaEventCreate(&e0), cudaEventCreate(&e1), cudaEventCreate(&f0), cudaEventCreate(&f1);
	for (uint32_t gl : {0u, 16u}) {
		for (int mode = 0; mode < 3; mode++) {  // 0 kernel, 1 copies, 2 both
			for (int rep = 0; rep < 2; rep++) {
				cudaDeviceSynchronize();
				auto t0 = std::chrono::steady_clock::now();
				if (mode != 1) {
					cudaEventRecord(e0, sa);
					k_spin<<<148, 512, 200 * 1024, sa>>>(dout, din, 60000, gl);
					cudaEventRecord(e1, sa);
				}
				if (mode != 0) {
					cudaEventRecord(f0, sb);
					for (size_t c = 0; c < NCH; c++) cudaMemcpyAsync(d + c * CH, h + c * CH, CH, cudaMemcpyHostToDevice, sb);
					cudaEventRecord(f1, sb);
				}
				cudaDeviceSynchronize();
				double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
				float km = 0, cm = 0;
				if (mode != 1) cudaEventElapsedTime(&km, e0, e1);
				if (mode != 0) cudaEventElapsedTime(&cm, f0, f1);
				if (rep) printf("global loads every %2u: mode %d  wall %.3f ms  kernel %.3f ms  copies %.3f ms\n", gl, mode, wall, km, cm);
			}
		}
	}
	return 0;
}
