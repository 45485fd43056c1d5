// ntt_lut.cuh -- additive-NTT pass kernel for B32 on nibble look-up tables (the fast path for every
// layer whose twiddles are shared by >= 32 butterflies).
//
// B200 has no GF(2^k) multiplier.  A butterfly multiplies a data element v by a twiddle t that is
// CONSTANT over 2^(i + log_x) butterflies of layer i, and v -> t*v is GF(2)-linear, so
//     t*v = XOR_{p<8} TAB_t[p][nibble_p(v)],      TAB_t[p][e] = t * (e << 4p)      (8 x 16 words per twiddle)
// i.e. 8 shared-memory gathers + ~17 ALU instructions per B32 product on the LSU pipe, against ~34
// ALU-pipe operations of the bit-sliced Karatsuba circuit (ntt_bs.cuh), which stays in use for the
// lowest layers (twiddles shared by < 32 butterflies) only.
//
// One pass = R = R1 + 3 layers (3 <= R <= 6) on tiles of 2^R rows x CC columns staged in shared memory
// by TMA bulk copies (double buffered), processed in two register-resident radix stages:
//   stage A: local layers 3..R-1, 2^R1 rows per thread; every lane of a warp uses the same twiddle
//   stage B: local layers 0..2,   8 rows per thread;   the lanes of a warp span the 2^R1 row groups
// Bank conflicts: a table block is [entry e:16][slot:4][position:8] words (entry stride 256 B, two blocks
// interleaved); a lane owns one slot (a replica in stage A, the twiddle of its row group in stage B) and
// walks the 8 nibble positions ROTATED by its index inside the slot (the data word is rotated once,
// 1 SHF), so the 32 lanes of every LDS hit 32 distinct banks by construction.  The address of a gather
// is one PRMT: (nibble << 8) | lane_byte, the table-set offset rides in the LDS immediate.
// Tables are GF(2)-linear in the twiddle and twiddles are GF(2)-linear in the block index
// (twiddle.rs:163-168), so a tile's tables are built from 32-word "basis products" s_evals[row][b] * 2^k
// (precomputed once per NTT object) by XORs only: ~450 cycles per tile id, amortised over all column
// chunks of that tile id (only the lowest pass rebuilds per tile).
//
// Reference semantics: crates/ntt/src/tests/reference.rs:68-160, single_threaded.rs:134-362 (see ntt.cuh).
#pragma once
#include "fold_tma.cuh"

namespace b200 {
namespace nttl {

constexpr uint32_t CWARPS = 8;                   // compute warps
constexpr uint32_t CTHREADS = CWARPS * 32;
constexpr uint32_t THREADS = CTHREADS + 32;      // + one IO warp (TMA loads / stores)
constexpr uint32_t NBUF = 2;                     // tile ring

struct Args {
	uint32_t *data;
	const uint32_t *basis;  // [32 rows][32 index bits][32]: s_evals[row][bit] * 2^k
	uint32_t lx, log_y;     // lx includes the extension-degree shift
	uint32_t i_lo;          // layers [i_lo, i_lo + R1 + 3)
	uint32_t row0, d;
	uint32_t log_cc;  // log2(columns per work item), 5..7
	uint32_t n_z;
	uint32_t nbits;   // index bits of the pass's lowest layer: log_y - 1 - i_lo + coset_bits
	uint64_t coset;
};

// byte offset of a table set inside the table area.  Stage B has 7 sets (heap index hB over local layers
// 2,1,0), stage A 2^R1 - 1 (heap index hA).  R1 = 3: a stage-B set holds 8 twiddles = a whole 4 KiB double
// block; otherwise a set is one 2 KiB block and two sets interleave in a double block.
__host__ __device__ constexpr uint32_t set_off(int R1, bool stage_a, uint32_t h) {
	if (R1 == 3) return stage_a ? 7u * 4096u + (h >> 1) * 4096u + (h & 1u) * 128u : h * 4096u;
	const uint32_t n = stage_a ? 7u + h : h;
	return (n >> 1) * 4096u + (n & 1u) * 128u;
}
__host__ __device__ constexpr uint32_t tab_bytes(int R1) { return R1 == 3 ? 11u * 4096u : 5u * 4096u; }

// Shared memory: tables | basis products of the pass's layers | per-twiddle and per-tile parts | mbarriers | tiles.
// Tile rows are padded so that BOTH register stages read and write it without bank conflicts: a warp owns
// WC = 32 >> R1 columns; row r = (h, r_lo) sits at h * hp + r_lo * p words with p = CC + WC, hp = 8 p + WC, so the
// lanes of stage A (8 r_lo x WC columns, h fixed) and of stage B (2^R1 h x WC columns, r_lo fixed) hit 32 banks.
struct Layout {
	uint32_t msm, bj, bt, bar, tile0, tile_bytes, p, hp, total;
};
__host__ __device__ inline Layout layout(int R1, uint32_t log_cc, uint32_t nbits) {
	const uint32_t R = R1 + 3, G = 1u << R1, CC = 1u << log_cc, WC = 32u >> R1;
	Layout L;
	L.msm = tab_bytes(R1);             // [R][nbits][32]
	L.bj = L.msm + R * nbits * 128u;   // [2^R][32] tile-independent part per twiddle (heap order)
	L.bt = L.bj + (128u << R);         // [R][32] tile part per layer
	L.bar = L.bt + R * 128u;           // full[NBUF], done[NBUF]
	L.tile0 = (L.bar + 8u * 2u * NBUF + 127u) & ~127u;
	L.p = CC + WC;
	L.hp = 8u * L.p + WC;
	L.tile_bytes = (G * L.hp * 4u + 127u) & ~127u;
	L.total = L.tile0 + NBUF * L.tile_bytes;
	return L;
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
	uint32_t r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
	return r;
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(ft_smem_u32(src)), "r"(bytes) : "memory");
}

// t * v for the twiddle(s) of the table set at byte offset `off` (a compile-time constant after unrolling)
__device__ __forceinline__ uint32_t lut_mul(const uint8_t *tab, uint32_t v, const uint32_t (&lb)[8], uint32_t rot4, uint32_t off) {
	const uint32_t w = __funnelshift_r(v, v, rot4);
	const uint32_t we = w & 0x0F0F0F0Fu, wo = (w >> 4) & 0x0F0F0F0Fu;
	uint32_t r[8];
#pragma unroll
	for (int k = 0; k < 8; k++) {
		// bytes: [0] = lane byte (slot, rotated position), [1] = nibble k of w, [2..3] = 0
		const uint32_t a = prmt((k & 1) ? wo : we, lb[k], 0x7604u | ((uint32_t)(k >> 1) << 4));
		r[k] = *reinterpret_cast<const uint32_t *>(tab + a + off);
	}
	return (r[0] ^ r[1] ^ r[2]) ^ (r[3] ^ r[4] ^ r[5]) ^ (r[6] ^ r[7]);
}

// LL layers on 2^LL register-resident rows; local layer m pairs x[q], x[q | 1 << m]; its twiddle set is
// heap(m, q >> (m + 1)).  Forward (reference.rs:92-109): u += v*t, v += u, layers descending;
// inverse (:140-157): v += u, u += v*t, ascending.
template <int LL, int R1, bool IS_A, bool INV>
__device__ __forceinline__ void run_layers(uint32_t (&x)[1 << LL], const uint8_t *tab, const uint32_t (&lb)[8], uint32_t rot4) {
#pragma unroll
	for (int step = 0; step < LL; step++) {
		const int m = INV ? step : LL - 1 - step;
#pragma unroll
		for (int q = 0; q < (1 << LL); q++) {
			if (q & (1 << m)) continue;
			const uint32_t s = (uint32_t)q >> (m + 1);
			const uint32_t off = set_off(R1, IS_A, (1u << (LL - 1 - m)) + s - 1u);
			uint32_t &u = x[q], &v = x[q | (1 << m)];
			if (!INV) {
				u ^= lut_mul(tab, v, lb, rot4, off);
				v ^= u;
			} else {
				v ^= u;
				u ^= lut_mul(tab, v, lb, rot4, off);
			}
		}
	}
}

__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(CTHREADS) : "memory"); }  // compute warps only
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ft_smem_u32(bar)) : "memory");
}

// Warp-specialised: the IO warp streams tiles (2^R rows x CC columns, one bulk copy per row) through a ring of
// NBUF buffers; the 8 compute warps each own strips of WC = 32 >> R1 columns of every tile and run both register
// stages on them with only __syncwarp in between (the exchange between the stages stays inside a column), so in
// the steady state no warp waits for another one; the compute warps meet (named barrier) only to rebuild the
// tables when the tile id changes.
template <int R1, bool INV>
__global__ void __launch_bounds__(THREADS, 2) k_ntt_lut(const Args A) {
	constexpr int R = R1 + 3;
	constexpr uint32_t G = 1u << R1, WC = 32u >> R1;
	extern __shared__ __align__(128) uint8_t smem[];
	const Layout L = layout(R1, A.log_cc, A.nbits);
	uint32_t *msm = reinterpret_cast<uint32_t *>(smem + L.msm);
	uint32_t *bj = reinterpret_cast<uint32_t *>(smem + L.bj);
	uint32_t *bt = reinterpret_cast<uint32_t *>(smem + L.bt);
	uint64_t *full = reinterpret_cast<uint64_t *>(smem + L.bar), *done = full + NBUF;
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t log_cc = A.log_cc, CC = 1u << log_cc, P = L.p, HP = L.hp, nbits = A.nbits;
	const uint32_t w = A.lx + A.i_lo;             // log2(columns of a tile id)
	const uint32_t log_chunks = w - log_cc;        // column chunks per tile id
	const uint32_t log_tiles = A.log_y - A.i_lo - R;
	const uint64_t total = (uint64_t)A.n_z << (log_tiles + log_chunks);
	const uint64_t it0 = total * blockIdx.x / gridDim.x, it1 = total * (blockIdx.x + 1) / gridDim.x;
	if (it0 >= it1) return;
	const uint32_t n_items = (uint32_t)(it1 - it0);

	if (tid == 0) {
		for (uint32_t b = 0; b < NBUF; b++) {
			ft_mbar_init(&full[b], 1);
			ft_mbar_init(&done[b], CWARPS);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();

	auto decode = [&](uint64_t item, uint32_t &z, uint32_t &T, uint32_t &chunk) {
		chunk = (uint32_t)(item & ((1ull << log_chunks) - 1));
		const uint64_t zt = item >> log_chunks;
		T = (uint32_t)(zt & ((1ull << log_tiles) - 1));
		z = (uint32_t)(zt >> log_tiles);
	};

	if (warp == CWARPS) {
		// ---- IO warp: store(item n - NBUF) then load(item n) through buffer n % NBUF ------------------------
		auto tile_rows = [&](uint64_t item) -> uint32_t * {
			uint32_t z, T, chunk;
			decode(item, z, T, chunk);
			return A.data + ((uint64_t)z << (A.lx + A.log_y)) + ((uint64_t)T << (R + w)) + ((uint64_t)chunk << log_cc);
		};
		for (uint32_t n = 0; n < n_items + NBUF; n++) {
			const uint32_t b = n % NBUF;
			uint8_t *buf = smem + L.tile0 + b * L.tile_bytes;
			if (n >= NBUF) {
				ft_mbar_wait(&done[b], ((n - NBUF) / NBUF) & 1u);
				uint32_t *dstg = tile_rows(it0 + n - NBUF);
				for (uint32_t r = lane; r < (1u << R); r += 32) bulk_s2g(dstg + ((uint64_t)r << w), buf + ((r >> 3) * HP + (r & 7u) * P) * 4u, CC * 4u);
				asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			}
			if (n < n_items) {
				if (n >= NBUF) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
				__syncwarp();
				const uint32_t *src = tile_rows(it0 + n);
				if (lane == 0) ft_mbar_expect_tx(&full[b], (CC * 4u) << R);
				__syncwarp();
				for (uint32_t r = lane; r < (1u << R); r += 32) ft_bulk_g2s(buf + ((r >> 3) * HP + (r & 7u) * P) * 4u, src + ((uint64_t)r << w), CC * 4u, &full[b]);
			}
		}
		asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
		return;
	}

	// ---- compute warps ------------------------------------------------------------------------------------
	// basis products of this pass's layers (rows row0 + i_lo ..): global -> shared, [R][nbits][32]
	for (uint32_t i = tid; i < R * nbits * 8u; i += CTHREADS) {
		const uint32_t l = i / (nbits * 8u), rem = i - l * nbits * 8u;
		reinterpret_cast<uint4 *>(msm)[i] = __ldg(reinterpret_cast<const uint4 *>(A.basis + (uint64_t)(A.row0 + A.i_lo + l) * 1024u) + rem);
	}
	cbar();
	// tile-independent part of every twiddle's basis products: heap index te = 2^(R-1-l) + jr
	for (uint32_t e = tid; e < (32u << R); e += CTHREADS) {
		const uint32_t te = e >> 5, b = e & 31u;
		uint32_t acc = 0;
		if (te) {
			const uint32_t lev = 31u - __clz(te), l = R - 1u - lev;
			uint32_t jr = te - (1u << lev);
			const uint32_t *mrow = msm + l * nbits * 32u + b;
			for (uint32_t bp = 0; jr; bp++, jr >>= 1)
				if (jr & 1u) acc ^= mrow[bp * 32u];
		}
		bj[e] = acc;
	}

	// lane constants: lane byte = slot * 32 + rotated position * 4 (+ 128 for the second block of a stage-B set)
	uint32_t lbA[8], lbB[8];
	uint32_t rotA4, rotB4;
	{
		const uint32_t slot = lane >> 3, rot = lane & 7u;
		rotA4 = 4u * rot;
#pragma unroll
		for (uint32_t k = 0; k < 8; k++) lbA[k] = slot * 32u + (((k + rot) & 7u) << 2);
		if (R1 == 3) {
			const uint32_t h = lane >> 2, blk = h >> 2, rotb = (lane & 3u) | (blk << 2);
			rotB4 = 4u * rotb;
#pragma unroll
			for (uint32_t k = 0; k < 8; k++) lbB[k] = blk * 128u + (h & 3u) * 32u + (((k + rotb) & 7u) << 2);
		} else {
			rotB4 = rotA4;
#pragma unroll
			for (uint32_t k = 0; k < 8; k++) lbB[k] = lbA[k];
		}
	}
	// the lane's place inside a strip of WC columns: stage B = (row group h, column a); stage A item j of the lane =
	// (r_lo, column) with 8 * WC items per strip
	const uint32_t hB = lane >> (5 - R1), aB = lane & (WC - 1u);
	const uint32_t n_strips = CC / WC;

	uint32_t cur_T = 0xFFFFFFFFu;
	for (uint32_t n = 0; n < n_items; n++) {
		const uint32_t buf = n % NBUF;
		uint32_t z, T, chunk;
		decode(it0 + n, z, T, chunk);
		if (T != cur_T) {
			cur_T = T;
			cbar();  // everybody is done with the old tables (and bj is complete the first time)
			// tile part of the basis products: idx = coset << (log_y-1-i) | T << (R-1-l)
			for (uint32_t e = tid; e < R * 32u; e += CTHREADS) {
				const uint32_t l = e >> 5, b = e & 31u, i = A.i_lo + l;
				uint64_t idx = (A.coset << (A.log_y - 1u - i)) | ((uint64_t)T << (R - 1u - l));
				const uint32_t *mrow = msm + l * nbits * 32u + b;
				uint32_t acc = 0;
				for (uint32_t bp = 0; idx && bp < nbits; bp++, idx >>= 1)
					if (idx & 1u) acc ^= mrow[bp * 32u];
				bt[e] = acc;
			}
			cbar();
			// tables: one warp-task per 2 KiB block; lane = (slot, position) builds the 16 entries of its twiddle
			constexpr uint32_t nB = (G == 8 ? 14u : 7u), n_tasks = nB + G - 1u;
			for (uint32_t task = warp; task < n_tasks; task += CWARPS) {
				const uint32_t slot = lane >> 3, pos = lane & 7u;
				uint32_t off, te, l;
				if (task < nB) {
					const uint32_t hb = G == 8 ? task >> 1 : task, blk = G == 8 ? task & 1u : 0u;
					const uint32_t lev = 31u - __clz(hb + 1u), m = 2u - lev, s = hb + 1u - (1u << lev);
					const uint32_t h = G == 8 ? blk * 4u + slot : slot >> (2 - (R1 < 3 ? R1 : 2));
					l = m;
					te = (1u << (R - 1u - m)) + ((h << (2u - m)) | s);
					off = set_off(R1, false, hb) + blk * 128u;
				} else {
					const uint32_t hA = task - nB;
					te = hA + 1u;
					l = 3u + (R1 - 1u - (31u - __clz(te)));
					off = set_off(R1, true, hA);
				}
				const uint4 p = reinterpret_cast<const uint4 *>(bt)[l * 8u + pos], q = reinterpret_cast<const uint4 *>(bj)[te * 8u + pos];
				const uint32_t b0 = p.x ^ q.x, b1 = p.y ^ q.y, b2 = p.z ^ q.z, b3 = p.w ^ q.w;
				uint32_t *dst = reinterpret_cast<uint32_t *>(smem + off) + slot * 8u + pos;
				const uint32_t b01 = b0 ^ b1, b23 = b2 ^ b3;
				dst[0 * 64] = 0;
				dst[1 * 64] = b0;
				dst[2 * 64] = b1;
				dst[3 * 64] = b01;
				dst[4 * 64] = b2;
				dst[5 * 64] = b2 ^ b0;
				dst[6 * 64] = b2 ^ b1;
				dst[7 * 64] = b2 ^ b01;
				dst[8 * 64] = b3;
				dst[9 * 64] = b3 ^ b0;
				dst[10 * 64] = b3 ^ b1;
				dst[11 * 64] = b3 ^ b01;
				dst[12 * 64] = b23;
				dst[13 * 64] = b23 ^ b0;
				dst[14 * 64] = b23 ^ b1;
				dst[15 * 64] = b23 ^ b01;
			}
			cbar();
		}
		ft_mbar_wait(&full[buf], (n / NBUF) & 1u);
		uint32_t *tile = reinterpret_cast<uint32_t *>(smem + L.tile0 + buf * L.tile_bytes);

		for (uint32_t strip = warp; strip < n_strips; strip += CWARPS) {
			uint32_t *sp = tile + strip * WC;
			auto stage_a = [&]() {
				if (R1 == 0) return;
				constexpr int LA = R1 > 0 ? R1 : 1;
#pragma unroll
				for (uint32_t j = 0; j < 8u / G; j++) {
					const uint32_t ia = lane + 32u * j;  // (r_lo, column): r_lo = ia / WC
					uint32_t *p = sp + (ia / WC) * P + (ia & (WC - 1u));
					uint32_t x[1 << LA];
#pragma unroll
					for (int q = 0; q < (1 << LA); q++) x[q] = p[q * HP];
					run_layers<LA, R1, true, INV>(x, smem, lbA, rotA4);
#pragma unroll
					for (int q = 0; q < (1 << LA); q++) p[q * HP] = x[q];
				}
			};
			auto stage_b = [&]() {
				uint32_t *p = sp + hB * HP + aB;
				uint32_t x[8];
#pragma unroll
				for (int q = 0; q < 8; q++) x[q] = p[q * P];
				run_layers<3, R1, false, INV>(x, smem, lbB, rotB4);
#pragma unroll
				for (int q = 0; q < 8; q++) p[q * P] = x[q];
			};
			if (!INV) {
				stage_a();
				if (R1 > 0) __syncwarp();
				stage_b();
			} else {
				stage_b();
				if (R1 > 0) __syncwarp();
				stage_a();
			}
		}
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		__syncwarp();
		if (lane == 0) mbar_arrive(&done[buf]);
	}
}

}  // namespace nttl
}  // namespace b200
