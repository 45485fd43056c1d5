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

// Launch geometry (picked by the host per pass): `cw` compute warps + two IO warps (TMA loader, TMA storer), a ring
// of `nbuf` tiles.  Passes whose tables live for many tiles run one CTA per SM with 16 compute warps and a deep
// ring (warps may drift apart by nbuf - 1 tiles); the lowest pass rebuilds its tables for every tile, which is
// a CTA-wide rendezvous, so it runs two CTAs of 8 compute warps per SM that fill each other's gaps.
constexpr uint32_t MAX_CW = 16;
constexpr uint32_t MAX_THREADS = MAX_CW * 32 + 64;
constexpr uint32_t MAX_NBUF = 4;

struct Args {
	uint32_t *data;
	const uint32_t *basis;  // [32 rows][32 index bits][32]: s_evals[row][bit] * 2^k
	uint32_t lx, log_y;     // lx includes the extension-degree shift
	uint32_t i_lo;          // layers [i_lo, i_lo + R1 + 3)
	uint32_t row0, d;
	uint32_t log_cc;  // log2(columns per work item), 5..7
	uint32_t n_z;
	uint32_t nbits;   // index bits of the pass's lowest layer: log_y - 1 - i_lo + coset_bits
	uint32_t cw, nbuf;  // compute warps (blockDim.x = 32 * (cw + 2)), tile ring depth
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
constexpr uint32_t BYTE_TAB_BYTES = 65536;  // [entry:256][set:2][slot:8][position:4] words: stage-A top layer | stage-B layer 2
__host__ __device__ inline Layout layout(int R1, uint32_t log_cc, uint32_t nbits, uint32_t nbuf, bool byte_tabs) {
	const uint32_t R = R1 + 3, G = 1u << R1, CC = 1u << log_cc, WC = 32u >> R1;
	Layout L;
	L.msm = tab_bytes(R1) + (byte_tabs ? BYTE_TAB_BYTES : 0u);  // [R][nbits][32]
	L.bj = L.msm + R * nbits * 128u;   // [2^R][32] tile-independent part per twiddle (heap order)
	L.bt = L.bj + (128u << R);         // [R][32] tile part per layer
	L.bar = L.bt + R * 128u;           // full, done, empty [MAX_NBUF]
	L.tile0 = (L.bar + 8u * 3u * MAX_NBUF + 127u) & ~127u;
	L.p = CC + WC;
	L.hp = 8u * L.p + WC;
	L.tile_bytes = (G * L.hp * 4u + 127u) & ~127u;
	L.total = L.tile0 + nbuf * L.tile_bytes;
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

// per-lane constants of one register stage
struct LaneC {
	uint32_t lb[8];   // nibble tables: lane byte = slot * 32 + rotated position * 4 (+ 128: second block of a stage-B set)
	uint32_t m1, m2;  // v * m1 / v * m2 (64-bit) hold rotr(v, 4 rot) / rotr(v, 4 rot + 4) as lo | hi
	uint32_t lbb[4];  // byte tables: lane byte = set * 128 + slot * 16 + rotated position * 4
	uint32_t mb;      // v * mb holds rotr(v, 8 rotb)
};
__device__ __forceinline__ uint32_t rot_mul(uint32_t s) { return (s & 31u) ? 1u << (32u - s) : 1u; }  // s in [0, 32]
// v * m as a 64-bit product on the FMA pipe (inline PTX: written in C the compiler sees the power of two and
// strength-reduces the product back into two ALU-pipe shifts)
__device__ __forceinline__ uint64_t mul_wide(uint32_t v, uint32_t m) {
	uint64_t p;
	asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(v), "r"(m));
	return p;
}
__device__ __forceinline__ uint32_t or_and(uint32_t a, uint32_t b, uint32_t c) {
	uint32_t r;
	asm("lop3.b32 %0, %1, %2, %3, 0xA8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));  // (a | b) & c
	return r;
}

// t * v for the twiddle(s) of the nibble-table set at byte offset `off` (a compile-time constant after unrolling).
// The rotation of v runs on the FMA pipe (IMAD.WIDE by a power of two), the OR of the two product halves and
// the nibble mask are one LOP3: 2 + 8 (PRMT) + 5 (XOR tree incl. the butterfly) ALU-pipe instructions per product.
__device__ __forceinline__ uint32_t lut_mul(const uint8_t *tab, uint32_t v, const LaneC &C, uint32_t off) {
	const uint64_t p1 = mul_wide(v, C.m1), p2 = mul_wide(v, C.m2);
	const uint32_t we = or_and((uint32_t)p1, (uint32_t)(p1 >> 32), 0x0F0F0F0Fu), wo = or_and((uint32_t)p2, (uint32_t)(p2 >> 32), 0x0F0F0F0Fu);
	uint32_t r[8];
#pragma unroll
	for (int k = 0; k < 8; k++) {
		// bytes: [0] = lane byte (slot, rotated position), [1] = nibble k of the rotated word, [2..3] = 0
		const uint32_t a = prmt((k & 1) ? wo : we, C.lb[k], 0x7604u | ((uint32_t)(k >> 1) << 4));
		r[k] = *reinterpret_cast<const uint32_t *>(tab + a + off);
	}
	return (r[0] ^ r[1] ^ r[2]) ^ (r[3] ^ r[4] ^ r[5]) ^ (r[6] ^ r[7]);
}
// the same product from a byte-table set (4 gathers): sets whose tables live for many tiles
__device__ __forceinline__ uint32_t lut_mul_byte(const uint8_t *tab, uint32_t v, const LaneC &C, uint32_t off) {
	const uint64_t p = mul_wide(v, C.mb);
	const uint32_t w = (uint32_t)p | (uint32_t)(p >> 32);
	uint32_t r[4];
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const uint32_t a = prmt(w, C.lbb[k], 0x7604u | ((uint32_t)k << 4));
		r[k] = *reinterpret_cast<const uint32_t *>(tab + a + off);
	}
	return (r[0] ^ r[1]) ^ (r[2] ^ r[3]);
}

// LL layers on 2^LL register-resident rows; local layer m pairs x[q], x[q | 1 << m]; its twiddle set is
// heap(m, q >> (m + 1)).  Forward (reference.rs:92-109): u += v*t, v += u, layers descending;
// inverse (:140-157): v += u, u += v*t, ascending.  BYTE: the top layer of the stage reads the byte-table set.
template <int LL, int R1, bool IS_A, bool INV, bool BYTE>
__device__ __forceinline__ void run_layers(uint32_t (&x)[1 << LL], const uint8_t *tab, const LaneC &C) {
#pragma unroll
	for (int step = 0; step < LL; step++) {
		const int m = INV ? step : LL - 1 - step;
#pragma unroll
		for (int q = 0; q < (1 << LL); q++) {
			if (q & (1 << m)) continue;
			const uint32_t s = (uint32_t)q >> (m + 1);
			const uint32_t off = set_off(R1, IS_A, (1u << (LL - 1 - m)) + s - 1u);
			uint32_t &u = x[q], &v = x[q | (1 << m)];
			if (INV) v ^= u;
			if (BYTE && m == LL - 1) u ^= lut_mul_byte(tab, v, C, tab_bytes(R1));
			else u ^= lut_mul(tab, v, C, off);
			if (!INV) v ^= u;
		}
	}
}

__device__ __forceinline__ void cbar(uint32_t n_threads) { asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory"); }  // compute warps only
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ft_smem_u32(bar)) : "memory");
}

// Warp-specialised: the IO warp streams tiles (2^R rows x CC columns, one bulk copy per row) through a ring of
// NBUF buffers; the 8 compute warps each own strips of WC = 32 >> R1 columns of every tile and run both register
// stages on them with only __syncwarp in between (the exchange between the stages stays inside a column), so in
// the steady state no warp waits for another one; the compute warps meet (named barrier) only to rebuild the
// tables when the tile id changes.
template <int R1, bool INV, bool BYTE>
__global__ void __launch_bounds__(MAX_THREADS, 1) k_ntt_lut(const Args A) {
	constexpr int R = R1 + 3;
	constexpr uint32_t G = 1u << R1, WC = 32u >> R1;
	extern __shared__ __align__(128) uint8_t smem[];
	const Layout L = layout(R1, A.log_cc, A.nbits, A.nbuf, BYTE);
	const uint32_t CWARPS = A.cw, CTHREADS = 32u * A.cw, NBUF = A.nbuf;
	uint32_t *msm = reinterpret_cast<uint32_t *>(smem + L.msm);
	uint32_t *bj = reinterpret_cast<uint32_t *>(smem + L.bj);
	uint32_t *bt = reinterpret_cast<uint32_t *>(smem + L.bt);
	uint64_t *full = reinterpret_cast<uint64_t *>(smem + L.bar), *done = full + MAX_NBUF, *empty = done + MAX_NBUF;
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
			ft_mbar_init(&empty[b], 1);
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

	if (warp >= CWARPS) {
		// ---- IO warps.  Loader: wait until buffer n % NBUF is empty, one bulk copy per tile row.  Storer: wait until
		// the compute warps are done with a tile, bulk-store its rows, and hand the buffer back once the copies have
		// READ it (the loader never waits behind a store of another buffer).
		auto tile_rows = [&](uint64_t item) -> uint32_t * {
			uint32_t z, T, chunk;
			decode(item, z, T, chunk);
			return A.data + ((uint64_t)z << (A.lx + A.log_y)) + ((uint64_t)T << (R + w)) + ((uint64_t)chunk << log_cc);
		};
		if (warp == CWARPS) {
			for (uint32_t n = 0; n < n_items; n++) {
				const uint32_t b = n % NBUF;
				uint8_t *buf = smem + L.tile0 + b * L.tile_bytes;
				if (n >= NBUF) ft_mbar_wait(&empty[b], ((n - NBUF) / NBUF) & 1u);
				const uint32_t *src = tile_rows(it0 + n);
				if (lane == 0) ft_mbar_expect_tx(&full[b], (CC * 4u) << R);
				__syncwarp();
				for (uint32_t r = lane; r < (1u << R); r += 32) ft_bulk_g2s(buf + ((r >> 3) * HP + (r & 7u) * P) * 4u, src + ((uint64_t)r << w), CC * 4u, &full[b]);
			}
		} else {
			for (uint32_t n = 0; n < n_items; n++) {
				const uint32_t b = n % NBUF;
				const uint8_t *buf = smem + L.tile0 + b * L.tile_bytes;
				ft_mbar_wait(&done[b], (n / NBUF) & 1u);
				uint32_t *dstg = tile_rows(it0 + n);
				for (uint32_t r = lane; r < (1u << R); r += 32) bulk_s2g(dstg + ((uint64_t)r << w), buf + ((r >> 3) * HP + (r & 7u) * P) * 4u, CC * 4u);
				asm volatile("cp.async.bulk.commit_group;" ::: "memory");
				asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
				__syncwarp();
				if (lane == 0) mbar_arrive(&empty[b]);
			}
			asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
		}
		return;
	}

	// ---- compute warps ------------------------------------------------------------------------------------
	// basis products of this pass's layers (rows row0 + i_lo ..): global -> shared, [R][nbits][32]
	for (uint32_t i = tid; i < R * nbits * 8u; i += CTHREADS) {
		const uint32_t l = i / (nbits * 8u), rem = i - l * nbits * 8u;
		reinterpret_cast<uint4 *>(msm)[i] = __ldg(reinterpret_cast<const uint4 *>(A.basis + (uint64_t)(A.row0 + A.i_lo + l) * 1024u) + rem);
	}
	cbar(CTHREADS);
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

	// lane constants of the two stages
	LaneC cA, cB;
	{
		const uint32_t slot = lane >> 3, rot = lane & 7u;
		cA.m1 = rot_mul(4u * rot), cA.m2 = rot_mul(4u * rot + 4u);
#pragma unroll
		for (uint32_t k = 0; k < 8; k++) cA.lb[k] = slot * 32u + (((k + rot) & 7u) << 2);
		if (R1 == 3) {
			const uint32_t h = lane >> 2, blk = h >> 2, rotb = (lane & 3u) | (blk << 2);
			cB.m1 = rot_mul(4u * rotb), cB.m2 = rot_mul(4u * rotb + 4u);
#pragma unroll
			for (uint32_t k = 0; k < 8; k++) cB.lb[k] = blk * 128u + (h & 3u) * 32u + (((k + rotb) & 7u) << 2);
		} else {
			cB.m1 = cA.m1, cB.m2 = cA.m2;
#pragma unroll
			for (uint32_t k = 0; k < 8; k++) cB.lb[k] = cA.lb[k];
		}
		// byte tables: slot = lane / 4 (a replica in stage A, (row group, replica) in stage B), position rotated by lane % 4
		const uint32_t bslot = lane >> 2, brot = lane & 3u;
		cA.mb = cB.mb = rot_mul(8u * brot);
#pragma unroll
		for (uint32_t k = 0; k < 4; k++) {
			cA.lbb[k] = bslot * 16u + (((k + brot) & 3u) << 2);
			cB.lbb[k] = 128u + cA.lbb[k];
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
			cbar(CTHREADS);  // everybody is done with the old tables (and bj is complete the first time)
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
			cbar(CTHREADS);
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
			if (BYTE) {
				// byte-table sets: 0 = top layer of stage A (one twiddle, 8 replicas), 1 = layer 2 of stage B (the
				// twiddle of the slot's row group); lane = (slot, position) builds its 256 entries from 8 basis products
				for (uint32_t set = (R1 > 0 ? 0u : 1u) + warp; set < 2u; set += CWARPS) {
					const uint32_t slot = lane >> 2, pos = lane & 3u;
					const uint32_t l = set == 0 ? R - 1u : 2u;
					const uint32_t te = set == 0 ? 1u : (1u << (R - 3u)) + (slot >> (3 - R1));
					const uint4 *pt = reinterpret_cast<const uint4 *>(bt) + l * 8u + 2u * pos, *pj = reinterpret_cast<const uint4 *>(bj) + te * 8u + 2u * pos;
					const uint4 lo4 = pt[0], hi4 = pt[1], lj = pj[0], hj = pj[1];
					const uint32_t b0 = lo4.x ^ lj.x, b1 = lo4.y ^ lj.y, b2 = lo4.z ^ lj.z, b3 = lo4.w ^ lj.w;
					const uint32_t b4 = hi4.x ^ hj.x, b5 = hi4.y ^ hj.y, b6 = hi4.z ^ hj.z, b7 = hi4.w ^ hj.w;
					uint32_t *dst = reinterpret_cast<uint32_t *>(smem + tab_bytes(R1)) + set * 32u + slot * 4u + pos;
					const uint32_t b01 = b0 ^ b1, b23 = b2 ^ b3;
					for (uint32_t hi = 0; hi < 16; hi++) {
						const uint32_t H = ((hi & 1u) ? b4 : 0u) ^ ((hi & 2u) ? b5 : 0u) ^ ((hi & 4u) ? b6 : 0u) ^ ((hi & 8u) ? b7 : 0u);
						uint32_t *d = dst + hi * 16u * 64u;
						d[0 * 64] = H;
						d[1 * 64] = H ^ b0;
						d[2 * 64] = H ^ b1;
						d[3 * 64] = H ^ b01;
						d[4 * 64] = H ^ b2;
						d[5 * 64] = H ^ b2 ^ b0;
						d[6 * 64] = H ^ b2 ^ b1;
						d[7 * 64] = H ^ b2 ^ b01;
						d[8 * 64] = H ^ b3;
						d[9 * 64] = H ^ b3 ^ b0;
						d[10 * 64] = H ^ b3 ^ b1;
						d[11 * 64] = H ^ b3 ^ b01;
						d[12 * 64] = H ^ b23;
						d[13 * 64] = H ^ b23 ^ b0;
						d[14 * 64] = H ^ b23 ^ b1;
						d[15 * 64] = H ^ b23 ^ b01;
					}
				}
			}
			cbar(CTHREADS);
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
					run_layers<LA, R1, true, INV, BYTE>(x, smem, cA);
#pragma unroll
					for (int q = 0; q < (1 << LA); q++) p[q * HP] = x[q];
				}
			};
			auto stage_b = [&]() {
				uint32_t *p = sp + hB * HP + aB;
				uint32_t x[8];
#pragma unroll
				for (int q = 0; q < 8; q++) x[q] = p[q * P];
				run_layers<3, R1, false, INV, BYTE>(x, smem, cB);
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
