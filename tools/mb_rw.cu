// mb_rw.cu -- memory-pattern microbenchmark for the fold kernel: read two 128 MiB streams, write one
// (in place or to a third buffer) with trivial arithmetic, to find what HBM delivers for the fold's
// 2:1 read:write mix.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mb_rw tools/mb_rw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int UNR, bool INPLACE, int HINT>
__global__ void __launch_bounds__(512) k_rw(uint4 *__restrict__ e0, const uint4 *__restrict__ e1, uint4 *__restrict__ out, uint64_t n) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * UNR;
	for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x * UNR + threadIdx.x; base < n; base += stride) {
		uint4 a[UNR], b[UNR];
#pragma unroll
		for (int u = 0; u < UNR; u++) {
			uint64_t i = base + (uint64_t)u * blockDim.x;
			if (i < n) {
				if (HINT == 1) {
					a[u] = __ldcs(e0 + i);
					b[u] = __ldcs(e1 + i);
				} else {
					a[u] = e0[i];
					b[u] = __ldg(e1 + i);
				}
			}
		}
#pragma unroll
		for (int u = 0; u < UNR; u++) {
			uint64_t i = base + (uint64_t)u * blockDim.x;
			if (i < n) {
				uint4 r = make_uint4(a[u].x ^ b[u].y, a[u].y ^ b[u].z, a[u].z ^ b[u].w, a[u].w ^ b[u].x);
				uint4 *dst = INPLACE ? e0 + i : out + i;
				if (HINT == 1) __stcs(dst, r);
				else *dst = r;
			}
		}
	}
}

__global__ void k_copy(const uint4 *__restrict__ a, uint4 *__restrict__ b, uint64_t n) {
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void k_read(const uint4 *__restrict__ a, uint4 *__restrict__ b, uint64_t n) {
	uint4 acc = make_uint4(0, 0, 0, 0);
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		uint4 v = a[i];
		acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
	}
	if (acc.x == 0x12345 && acc.y == 7) b[0] = acc;
}

template <typename F> float time_ms(F f, int reps = 20) {
	cudaEvent_t a, b;
	cudaEventCreate(&a);
	cudaEventCreate(&b);
	for (int i = 0; i < 3; i++) f();
	cudaEventRecord(a);
	for (int i = 0; i < reps; i++) f();
	cudaEventRecord(b);
	cudaEventSynchronize(b);
	float ms;
	cudaEventElapsedTime(&ms, a, b);
	return ms / reps;
}

int main() {
	const uint64_t n = 1ull << 23;  // elements per stream (128 MiB)
	uint4 *e0, *e1, *out;
	cudaMalloc(&e0, n * 16);
	cudaMalloc(&e1, n * 16);
	cudaMalloc(&out, n * 16);
	cudaMemset(e0, 1, n * 16);
	cudaMemset(e1, 2, n * 16);
	const double gb = 3.0 * n * 16 / 1e9;
	int sms = 148;
	for (int per_sm : {1, 2, 3, 4, 8, 16}) {
		int grid = sms * per_sm;
		float t;
		t = time_ms([&] { k_rw<2, true, 0><<<grid, 512>>>(e0, e1, out, n); });
		printf("grid %4d inplace unr2      : %.4f ms  %.0f GB/s\n", grid, t, gb / t * 1e3);
		t = time_ms([&] { k_rw<4, true, 0><<<grid, 512>>>(e0, e1, out, n); });
		printf("grid %4d inplace unr4      : %.4f ms  %.0f GB/s\n", grid, t, gb / t * 1e3);
		t = time_ms([&] { k_rw<2, false, 0><<<grid, 512>>>(e0, e1, out, n); });
		printf("grid %4d outofplace unr2   : %.4f ms  %.0f GB/s\n", grid, t, gb / t * 1e3);
		t = time_ms([&] { k_rw<2, true, 1><<<grid, 512>>>(e0, e1, out, n); });
		printf("grid %4d inplace unr2 cs   : %.4f ms  %.0f GB/s\n", grid, t, gb / t * 1e3);
		t = time_ms([&] { k_rw<4, false, 1><<<grid, 512>>>(e0, e1, out, n); });
		printf("grid %4d outofplace unr4 cs: %.4f ms  %.0f GB/s\n", grid, t, gb / t * 1e3);
	}
	{
		float t = time_ms([&] { k_copy<<<148 * 16, 512>>>(e0, out, 2 * n < n ? n : n); });
		printf("copy 128 MiB: %.4f ms %.0f GB/s (r+w)\n", t, 2.0 * n * 16 / 1e9 / t * 1e3);
		t = time_ms([&] { k_read<<<148 * 16, 512>>>(e0, out, n); });
		printf("read 128 MiB: %.4f ms %.0f GB/s\n", t, 1.0 * n * 16 / 1e9 / t * 1e3);
		t = time_ms([&] { cudaMemcpyAsync(out, e0, n * 16, cudaMemcpyDeviceToDevice); });
		printf("cudaMemcpy D2D 128 MiB: %.4f ms %.0f GB/s (r+w)\n", t, 2.0 * n * 16 / 1e9 / t * 1e3);
	}
	return 0;
}
