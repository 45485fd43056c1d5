// mb_pcie.cu -- what the PCIe link of the box delivers for the e2e path: pinned H2D, D2H, and both at once.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mb_pcie tools/mb_pcie.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <chrono>
int main() {
	const size_t N = 256u << 20;
	void *h0, *h1, *d0, *d1;
	cudaHostAlloc(&h0, N, cudaHostAllocDefault);
	cudaHostAlloc(&h1, N, cudaHostAllocDefault);
	cudaMalloc(&d0, N);
	cudaMalloc(&d1, N);
	cudaStream_t s0, s1;
	cudaStreamCreate(&s0);
	cudaStreamCreate(&s1);
	auto run = [&](const char *name, size_t chunk, bool h2d, bool d2h) {
		for (int rep = 0; rep < 2; rep++) {
			cudaDeviceSynchronize();
			auto t0 = std::chrono::steady_clock::now();
			for (size_t off = 0; off < N; off += chunk) {
				if (h2d) cudaMemcpyAsync((char *)d0 + off, (char *)h0 + off, chunk, cudaMemcpyHostToDevice, s0);
				if (d2h && off < N / 2) cudaMemcpyAsync((char *)h1 + off, (char *)d1 + off, chunk, cudaMemcpyDeviceToHost, s1);
			}
			cudaDeviceSynchronize();
			double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
			if (rep) printf("%-34s chunk %4zu MiB: %.3f ms  H2D %.1f GB/s\n", name, chunk >> 20, ms, h2d ? N / ms / 1e6 : (N / 2) / ms / 1e6);
		}
	};
	for (size_t c : {(size_t)16 << 20, (size_t)64 << 20, (size_t)256 << 20}) {
		run("H2D 256 MiB", c, true, false);
		run("D2H 128 MiB", c, false, true);
		run("H2D 256 MiB + D2H 128 MiB", c, true, true);
	}
	return 0;
}
