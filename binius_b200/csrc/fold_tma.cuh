// fold_tma.cuh -- the fold-high / fold-low kernel with TMA-staged operands.
//
// Same arithmetic as k_lerp_lut (kernels.cuh): e0[i] ^= (e1[i] ^ e0[i]) * z with one broadcast
// challenge z, served by the Karatsuba-64 byte-LUT engine (linmap.cuh).  What changes is how the
// operands reach the SM: one persistent CTA of 32 warps per SM; EVERY WARP keeps its own ring of NSTAGE
// 64-element tiles in shared memory, filled by `cp.async.bulk` (TMA 1-D bulk copies, SASS UBLKCP)
// that complete on per-warp mbarriers, so 128 KiB per SM are always in flight, no warp ever stalls on a
// global load and there is no CTA-wide barrier in the streaming loop (warps drift apart, which keeps
// the LDS, ALU and store pipes overlapped)
// (the register-staged kernel was latency-bound: ncu long_scoreboard 4.8 of 12 stall cycles per issue
// at 65% of HBM peak with the LDS pipe no longer the limiter).  Results go straight from registers to
// HBM with STG.128.
//
// reference semantics: compute/src/cpu/layer.rs:393-408 (extrapolate_line), math/src/fold.rs:648-696
// (fold_left_lerp_inplace, const suffix) and :528-575 (fold_right_lerp, PAIRS = true).
#pragma once
#include "kernels.cuh"

namespace b200 {

constexpr uint32_t FT_THREADS = 1024;
constexpr uint32_t FT_WARPS = FT_THREADS / 32;
constexpr uint32_t FT_UNR = 2;
constexpr uint32_t FT_TILE = 32 * FT_UNR;              // outputs per warp-tile: every warp streams its own tiles
constexpr uint32_t FT_NSTAGE = 2;
constexpr uint32_t FT_STAGE_BYTES = 2 * FT_TILE * 16;  // 2 KiB: [e0 tile | e1 tile] or 2*TILE interleaved inputs
constexpr uint32_t FT_SMEM = LUT_BYTES + 2048 + FT_WARPS * FT_NSTAGE * FT_STAGE_BYTES + FT_WARPS * FT_NSTAGE * 8;
constexpr uint32_t FT_MAX_SEGS = 448;  // segment lists beyond the by-value 48 are copied to shared memory (28 KiB)

__device__ __forceinline__ uint32_t ft_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ft_mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ft_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ft_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ft_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ft_mbar_wait(uint64_t *bar, uint32_t parity) {
	uint32_t done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
					 : "=r"(done)
					 : "r"(ft_smem_u32(bar)), "r"(parity)
					 : "memory");
	} while (!done);
}
__device__ __forceinline__ void ft_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ft_smem_u32(dst)), "l"(src),
				 "r"(bytes), "r"(ft_smem_u32(bar))
				 : "memory");
}

// what one warp-tile covers; computed once per tile, NSTAGE iterations before it is consumed.
// All sizes are 32-bit: the host only launches this kernel for segments below 2^32 elements.
struct FtDesc {
	uint32_t seg;   // segment index
	uint32_t base;  // first output index inside the segment
	uint32_t cnt;   // outputs in this tile (0: past the end)
	uint32_t cntb;  // outputs whose second operand is stored (the rest pair with the suffix)
};
// cursor over the segment list for an increasing tile sequence
struct FtCursor {
	uint32_t seg, seg_start, seg_end;
};

template <bool PAIRS, bool SMEM_SEGS>
__global__ void __launch_bounds__(FT_THREADS, 1) k_lerp_tma(const __grid_constant__ LerpArgs A) {
	extern __shared__ __align__(256) uint8_t smem[];
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint8_t *ring = smem + LUT_BYTES + 2048 + warp * (FT_NSTAGE * FT_STAGE_BYTES);
	uint64_t *full = reinterpret_cast<uint64_t *>(smem + LUT_BYTES + 2048 + FT_WARPS * FT_NSTAGE * FT_STAGE_BYTES) + warp * FT_NSTAGE;
	const uint32_t n_tiles = (uint32_t)A.n_tiles, tile_step = gridDim.x * FT_WARPS;
	uint32_t next_tile = blockIdx.x * FT_WARPS + warp;  // tile of the next descriptor to build

	// up to LERP_MAX_SEGS segments travel by value in the kernel parameters (SMEM_SEGS = false: read straight
	// from the constant bank); longer lists are staged in device memory and copied to shared memory once per
	// CTA (descriptor reads sit on the critical path of every tile request: never a global load)
	const LerpSeg *segs = A.segs;
	if (SMEM_SEGS) {
		uint4 *sseg = reinterpret_cast<uint4 *>(smem + FT_SMEM);
		const uint4 *g = reinterpret_cast<const uint4 *>(A.segs_dev);
		for (uint32_t i = threadIdx.x; i < A.n_segs * (sizeof(LerpSeg) / 16); i += blockDim.x) sseg[i] = __ldg(g + i);
		__syncthreads();
		segs = reinterpret_cast<const LerpSeg *>(sseg);
	}
#define FT_SEG(i) (SMEM_SEGS ? segs[i] : A.segs[i])
	FtCursor C{0, 0, A.n_segs > 1 ? (uint32_t)FT_SEG(1).tile_start : n_tiles};
	// build the descriptor of `next_tile` (warp-uniform), have lane 0 request it into stage `s`, advance
	auto request = [&](uint32_t s) -> FtDesc {
		FtDesc d{0, 0, 0, 0};
		if (next_tile < n_tiles) {
			while (next_tile >= C.seg_end) {
				C.seg++;
				C.seg_start = C.seg_end;
				C.seg_end = C.seg + 1 < A.n_segs ? (uint32_t)FT_SEG(C.seg + 1).tile_start : n_tiles;
			}
			const LerpSeg &S = FT_SEG(C.seg);
			d.seg = C.seg;
			d.base = (next_tile - C.seg_start) * FT_TILE;
			d.cnt = min(FT_TILE, (uint32_t)S.upper - d.base);
			const uint32_t pivot = (uint32_t)S.pivot;
			d.cntb = min(d.cnt, max(pivot, d.base) - d.base);
			if (lane == 0) {
				uint8_t *dst = ring + s * FT_STAGE_BYTES;
				ft_mbar_expect_tx(&full[s], (d.cnt + d.cntb) * 16);
				if (PAIRS) {
					ft_bulk_g2s(dst, S.e1 + 2 * (uint64_t)d.base, (d.cnt + d.cntb) * 16, &full[s]);  // interleaved (2i, 2i+1)
				} else {
					ft_bulk_g2s(dst, S.e0 + d.base, d.cnt * 16, &full[s]);
					if (d.cntb) ft_bulk_g2s(dst + FT_TILE * 16, S.e1 + d.base, d.cntb * 16, &full[s]);
				}
			}
			next_tile += tile_step;
		}
		return d;
	};
	if (lane == 0) {
		for (uint32_t s = 0; s < FT_NSTAGE; s++) ft_mbar_init(&full[s], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncwarp();
	static_assert(FT_NSTAGE == 2, "the descriptor ring below is written for two stages");
	FtDesc d0 = request(0), d1 = request(1);  // loads fly while the table is built
	const MulEngine<true> E(smem, A.z);       // ends with __syncthreads

	for (uint32_t k = 0; d0.cnt; k++) {
		const uint32_t s = k & 1u;
		const LerpSeg &S = FT_SEG(d0.seg);
		const uint4 *st = reinterpret_cast<const uint4 *>(ring + s * FT_STAGE_BYTES);
		ft_mbar_wait(&full[s], (k >> 1) & 1u);
		uint4 a[FT_UNR], x[FT_UNR];
#pragma unroll
		for (uint32_t u = 0; u < FT_UNR; u++) {
			const uint32_t i = lane + 32 * u;
			// lanes past the end read stale (finite) shared memory; their result is never stored
			a[u] = PAIRS ? st[2 * i] : st[i];
			uint4 b = PAIRS ? st[2 * i + 1] : st[FT_TILE + i];
			if (i >= d0.cntb) b = S.suffix;
			x[u] = a[u] ^ b;
		}
		// the XORs consumed the shared-memory loads: once the whole warp is here the stage may be refilled
		__syncwarp();
		const FtDesc d2 = request(s);
		uint4 *out = S.e0 + d0.base;
#pragma unroll
		for (uint32_t u = 0; u < FT_UNR; u++) x[u] = E.mul(x[u]);
#pragma unroll
		for (uint32_t u = 0; u < FT_UNR; u++) {
			const uint32_t i = lane + 32 * u;
			if (i < d0.cnt) out[i] = a[u] ^ x[u];
		}
		d0 = d1;
		d1 = d2;
	}
}

#undef FT_SEG

}  // namespace b200
