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

// what one tile covers
struct FtTile {
	uint64_t base;  // first output index inside the segment
	uint32_t cnt;   // outputs in this tile
	uint32_t cntb;  // outputs whose second operand is stored (the rest pair with the suffix)
};
__device__ __forceinline__ FtTile ft_tile(const LerpSeg &S, uint64_t tile) {
	FtTile t;
	t.base = (tile - S.tile_start) * FT_TILE;
	const uint64_t left = S.upper - t.base;
	t.cnt = left < FT_TILE ? (uint32_t)left : FT_TILE;
	const uint64_t lb = S.pivot > t.base ? S.pivot - t.base : 0;
	t.cntb = lb < t.cnt ? (uint32_t)lb : t.cnt;
	return t;
}

template <bool PAIRS>
__global__ void __launch_bounds__(FT_THREADS, 1) k_lerp_tma(const __grid_constant__ LerpArgs A) {
	extern __shared__ __align__(256) uint8_t smem[];
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint8_t *ring = smem + LUT_BYTES + 2048 + warp * (FT_NSTAGE * FT_STAGE_BYTES);
	uint64_t *full = reinterpret_cast<uint64_t *>(smem + LUT_BYTES + 2048 + FT_WARPS * FT_NSTAGE * FT_STAGE_BYTES) + warp * FT_NSTAGE;
	const uint64_t tile0 = (uint64_t)blockIdx.x * FT_WARPS + warp, tile_step = (uint64_t)gridDim.x * FT_WARPS;

	// producer state (lane 0 of every warp): next tile to request and its segment cursor
	uint64_t p_tile = tile0;
	uint32_t p_seg = 0, p_k = 0;
	auto issue = [&]() {
		if (p_tile >= A.n_tiles) return;
		while (p_seg + 1 < A.n_segs && A.segs[p_seg + 1].tile_start <= p_tile) p_seg++;
		const LerpSeg &S = A.segs[p_seg];
		const FtTile t = ft_tile(S, p_tile);
		const uint32_t s = p_k % FT_NSTAGE;
		uint8_t *dst = ring + s * FT_STAGE_BYTES;
		ft_mbar_expect_tx(&full[s], (t.cnt + t.cntb) * 16);
		if (PAIRS) {
			ft_bulk_g2s(dst, S.e1 + 2 * t.base, (t.cnt + t.cntb) * 16, &full[s]);  // interleaved (2i, 2i+1)
		} else {
			ft_bulk_g2s(dst, S.e0 + t.base, t.cnt * 16, &full[s]);
			if (t.cntb) ft_bulk_g2s(dst + FT_TILE * 16, S.e1 + t.base, t.cntb * 16, &full[s]);
		}
		p_tile += tile_step;
		p_k++;
	};
	if (lane == 0) {
		for (uint32_t s = 0; s < FT_NSTAGE; s++) ft_mbar_init(&full[s], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		for (uint32_t s = 0; s < FT_NSTAGE; s++) issue();  // loads fly while the table is built
	}
	const MulEngine<true> E(smem, A.z);  // ends with __syncthreads

	uint32_t seg = 0, k = 0;
	for (uint64_t tile = tile0; tile < A.n_tiles; tile += tile_step, k++) {
		while (seg + 1 < A.n_segs && A.segs[seg + 1].tile_start <= tile) seg++;
		const LerpSeg &S = A.segs[seg];
		const FtTile t = ft_tile(S, tile);
		const uint32_t s = k % FT_NSTAGE;
		const uint4 *st = reinterpret_cast<const uint4 *>(ring + s * FT_STAGE_BYTES);
		ft_mbar_wait(&full[s], (k / FT_NSTAGE) & 1u);
		uint4 a[FT_UNR], x[FT_UNR];
#pragma unroll
		for (uint32_t u = 0; u < FT_UNR; u++) {
			const uint32_t i = lane + 32 * u;
			a[u] = make_uint4(0, 0, 0, 0);
			x[u] = a[u];
			if (i < t.cnt) {
				a[u] = PAIRS ? st[2 * i] : st[i];
				const uint4 b = i < t.cntb ? (PAIRS ? st[2 * i + 1] : st[FT_TILE + i]) : S.suffix;
				x[u] = a[u] ^ b;
			}
		}
		// the XORs consumed the shared-memory loads: once the whole warp is here the stage may be refilled
		__syncwarp();
		if (lane == 0) issue();
#pragma unroll
		for (uint32_t u = 0; u < FT_UNR; u++) {
			const uint32_t i = lane + 32 * u;
			if (i < t.cnt) S.e0[t.base + i] = a[u] ^ E.mul(x[u]);
		}
	}
}

}  // namespace b200
