// linmap.cuh -- the shared-memory LUT engine for GF(2)-linear maps  L : {0,1}^128 -> B128.
//
// Multiplication of a B128 element x by a *broadcast constant* z (the sumcheck fold challenge, a
// tensor-expansion coordinate, a FRI fold challenge) is GF(2)-linear in the bits of x:
//     x * z = XOR_{k=0..15}  T_k[ byte_k(x) ],      T_k[b] = XOR_{i : bit i of b} (beta_{8k+i} * z)
// where beta_i = 1 << i is the tower basis (reference crates/field/src/binary_field.rs:600-607).
// B200 has neither GFNI nor CLMUL, and 1-bit MMA is emulated through IMMA on sm_100a, so the
// per-byte table (Four-Russians) form is the cheapest exact evaluation: 16 x LDS.128 + ~64 ALU ops
// per product, against ~5000 word-ops for a straight tower multiply.
//
// Shared-memory layout (64 KiB): entry (k, b) lives at byte offset
//       (k >> 3) * 32768  +  b * 128  +  (k & 7) * 16
// i.e. table k occupies ONE 16-byte bank-quad column.  Lane j (= lane & 7) of every quarter-warp
// walks the byte positions in the order k = s ^ j (s = 0..15), so in every LDS.128 the 8 lanes of a
// quarter-warp hit 8 different bank-quads: the data-dependent gathers are bank-conflict-free by
// construction.  The element's bytes are pre-permuted per lane (word swap + PRMT) so that the
// byte extraction itself stays a compile-time-constant PRMT.
#pragma once
#include "field.cuh"

namespace b200 {

constexpr uint32_t LUT_BYTES = 65536;

struct LutLane {
	uint32_t sel;     // PRMT selector that XOR-permutes the bytes of a word by (j & 3)
	uint32_t swap;    // nonzero: swap word pairs (j >> 2)
	uint32_t off[8];  // ((s & 7) ^ j) << 4
};

__device__ __forceinline__ LutLane lut_lane_init() {
	LutLane L;
	uint32_t j = threadIdx.x & 7;
	L.sel = 0x3210u ^ ((j & 3u) * 0x1111u);
	L.swap = (j >> 2) & 1u;
#pragma unroll
	for (uint32_t s = 0; s < 8; s++) L.off[s] = ((s ^ j) << 4);
	return L;
}

// Build the 64 KiB table from the 128 basis images W[i] = L(beta_i) (global memory).
// `stage` = 2 KiB of shared memory used to stage W.  All threads of the CTA must call.
__device__ __forceinline__ void lut_build(uint8_t *tbl, uint4 *stage, const uint4 *__restrict__ W) {
	for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) stage[i] = __ldg(W + i);
	__syncthreads();
	for (uint32_t e = threadIdx.x; e < 4096; e += blockDim.x) {
		uint32_t k = e & 15, b = e >> 4;
		uint4 acc = u4_zero();
#pragma unroll
		for (uint32_t i = 0; i < 8; i++) {
			uint32_t m = 0u - ((b >> i) & 1u);
			uint4 w = stage[8 * k + i];
			acc.x ^= w.x & m; acc.y ^= w.y & m; acc.z ^= w.z & m; acc.w ^= w.w & m;
		}
		*reinterpret_cast<uint4 *>(tbl + (k >> 3) * 32768 + b * 128 + (k & 7) * 16) = acc;
	}
	__syncthreads();
}

__device__ __forceinline__ uint4 lut_ld(const uint8_t *tbl, uint32_t w, uint32_t bytepos, uint32_t blk, uint32_t off) {
	uint32_t b = __byte_perm(w, 0u, 0x4440u | bytepos);  // zero-extended byte
	return *reinterpret_cast<const uint4 *>(tbl + blk * 32768u + b * 128u + off);
}

// y = L(x)
__device__ __forceinline__ uint4 lut_apply(const uint8_t *tbl, const LutLane &L, uint4 x) {
	uint32_t w0 = L.swap ? x.y : x.x, w1 = L.swap ? x.x : x.y;
	uint32_t w2 = L.swap ? x.w : x.z, w3 = L.swap ? x.z : x.w;
	w0 = __byte_perm(w0, 0u, L.sel);
	w1 = __byte_perm(w1, 0u, L.sel);
	w2 = __byte_perm(w2, 0u, L.sel);
	w3 = __byte_perm(w3, 0u, L.sel);
	uint4 a0 = lut_ld(tbl, w0, 0, 0, L.off[0]) ^ lut_ld(tbl, w0, 1, 0, L.off[1]);
	uint4 a1 = lut_ld(tbl, w0, 2, 0, L.off[2]) ^ lut_ld(tbl, w0, 3, 0, L.off[3]);
	uint4 a2 = lut_ld(tbl, w1, 0, 0, L.off[4]) ^ lut_ld(tbl, w1, 1, 0, L.off[5]);
	uint4 a3 = lut_ld(tbl, w1, 2, 0, L.off[6]) ^ lut_ld(tbl, w1, 3, 0, L.off[7]);
	a0 ^= lut_ld(tbl, w2, 0, 1, L.off[0]) ^ lut_ld(tbl, w2, 1, 1, L.off[1]);
	a1 ^= lut_ld(tbl, w2, 2, 1, L.off[2]) ^ lut_ld(tbl, w2, 3, 1, L.off[3]);
	a2 ^= lut_ld(tbl, w3, 0, 1, L.off[4]) ^ lut_ld(tbl, w3, 1, 1, L.off[5]);
	a3 ^= lut_ld(tbl, w3, 2, 1, L.off[6]) ^ lut_ld(tbl, w3, 3, 1, L.off[7]);
	return (a0 ^ a1) ^ (a2 ^ a3);
}

}  // namespace b200
