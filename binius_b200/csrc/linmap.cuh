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

// ---- multiplication by the tower generators, limb-wise (no tables) -------------------------------
// alpha_w<K>(w): multiply every 2^K-bit limb of the word w by X_{K-1}  (K <= 5), i.e. mul_alpha at
// level K (reference pairwise_recursive_arithmetic.rs:48-62): (a0, a1) -> (a1, a0 + alpha_{K-1}(a1)).
template <int K> __device__ __forceinline__ uint32_t alpha_w(uint32_t w) {
	constexpr uint32_t M = K == 1 ? 0x55555555u : K == 2 ? 0x33333333u : K == 3 ? 0x0F0F0F0Fu : K == 4 ? 0x00FF00FFu : 0x0000FFFFu;
	constexpr int H = 1 << (K - 1);
	uint32_t hi = (w >> H) & M, lo = w & M;
	return hi | ((lo ^ alpha_w<K - 1>(hi)) << H);
}
template <> __device__ __forceinline__ uint32_t alpha_w<0>(uint32_t w) { return w; }

// multiply a B128 element by X_k = beta_{2^k} (k = 0..6)
__device__ __forceinline__ uint4 mul_tower_gen(uint4 v, uint32_t k) {
	switch (k) {
	case 0: return make_uint4(alpha_w<1>(v.x), alpha_w<1>(v.y), alpha_w<1>(v.z), alpha_w<1>(v.w));
	case 1: return make_uint4(alpha_w<2>(v.x), alpha_w<2>(v.y), alpha_w<2>(v.z), alpha_w<2>(v.w));
	case 2: return make_uint4(alpha_w<3>(v.x), alpha_w<3>(v.y), alpha_w<3>(v.z), alpha_w<3>(v.w));
	case 3: return make_uint4(alpha_w<4>(v.x), alpha_w<4>(v.y), alpha_w<4>(v.z), alpha_w<4>(v.w));
	case 4: return make_uint4(alpha_w<5>(v.x), alpha_w<5>(v.y), alpha_w<5>(v.z), alpha_w<5>(v.w));
	case 5: return make_uint4(v.y, v.x ^ alpha_w<5>(v.y), v.w, v.z ^ alpha_w<5>(v.w));
	default: return make_uint4(v.z, v.w, v.x ^ v.w, v.y ^ v.z ^ alpha_w<5>(v.w));  // (lo,hi)->(hi, lo + alpha_6(hi))
	}
}

// beta_i * z for the tower basis beta_i = 1 << i = prod_{k : bit k of i} X_k
__device__ __forceinline__ uint4 basis_image(uint4 z, uint32_t i) {
#pragma unroll
	for (uint32_t k = 0; k < 7; k++)
		if ((i >> k) & 1u) z = mul_tower_gen(z, k);
	return z;
}

// Build the 64 KiB table of the map x -> x * z.  `stage` = 2 KiB of shared memory holding the basis
// images transposed (stage[bit * 16 + k] = beta_{8k+bit} * z) so that the 8 lanes of a quarter-warp
// read and write 8 different bank-quads (conflict-free build).  All threads of the CTA must call.
__device__ __forceinline__ void lut_build_images(uint8_t *tbl, const uint4 *stage) {
	for (uint32_t e = threadIdx.x; e < 4096; e += blockDim.x) {
		uint32_t k = e & 15, b = e >> 4;
		uint4 acc = u4_zero();
#pragma unroll
		for (uint32_t i = 0; i < 8; i++) {
			uint32_t m = 0u - ((b >> i) & 1u);
			uint4 w = stage[i * 16 + k];
			acc.x ^= w.x & m; acc.y ^= w.y & m; acc.z ^= w.z & m; acc.w ^= w.w & m;
		}
		*reinterpret_cast<uint4 *>(tbl + (k >> 3) * 32768 + b * 128 + (k & 7) * 16) = acc;
	}
	__syncthreads();
}
__device__ __forceinline__ void lut_build_mul(uint8_t *tbl, uint4 *stage, uint4 z) {
	for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) stage[(i & 7) * 16 + (i >> 3)] = basis_image(z, i);
	__syncthreads();
	lut_build_images(tbl, stage);
}
// general linear map given by 128 basis images in global memory (W[i] = L(beta_i))
__device__ __forceinline__ void lut_build(uint8_t *tbl, uint4 *stage, const uint4 *__restrict__ W) {
	for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) stage[(i & 7) * 16 + (i >> 3)] = __ldg(W + i);
	__syncthreads();
	lut_build_images(tbl, stage);
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

// ------------------------------------------------------------------------------------------------
// Nibble variant: 32 tables of 16 entries = 8 KiB per constant instead of 64 KiB, 32 gathers per
// product instead of 16.  For ops that multiply by SEVERAL constants in one kernel (fri_fold: one
// table per folding challenge).  Entry (p, e) of nibble position p lives at
//       (p >> 3) * 2048 + e * 128 + (p & 7) * 16
// and lane j (= lane & 7) visits the positions in the order p = s ^ j: same conflict-free argument
// as above.  The per-lane pre-permutation is a PRMT (bytes, bits 1-2 of j) plus a nibble swap (bit 0).
constexpr uint32_t NLUT_BYTES = 8192;

struct NLutLane {
	uint32_t sel;     // PRMT selector: byte index XOR ((j >> 1) & 3)
	uint32_t nswap;   // nonzero: swap the nibbles of every byte (j & 1)
	uint32_t off[8];  // ((s & 7) ^ j) << 4
};
__device__ __forceinline__ NLutLane nlut_lane_init() {
	NLutLane L;
	uint32_t j = threadIdx.x & 7;
	L.sel = 0x3210u ^ (((j >> 1) & 3u) * 0x1111u);
	L.nswap = j & 1u;
#pragma unroll
	for (uint32_t s = 0; s < 8; s++) L.off[s] = ((s ^ j) << 4);
	return L;
}
// build the table of x -> x * z; all threads of the CTA must call (ends with __syncthreads)
__device__ __forceinline__ void nlut_build_mul(uint8_t *tbl, uint4 z) {
	for (uint32_t e = threadIdx.x; e < 512; e += blockDim.x) {
		uint32_t p = e & 31, v = e >> 5;  // consecutive lanes -> consecutive positions -> distinct bank-quads
		uint4 acc = u4_zero();
#pragma unroll
		for (uint32_t i = 0; i < 4; i++)
			if ((v >> i) & 1u) acc ^= basis_image(z, 4 * p + i);
		*reinterpret_cast<uint4 *>(tbl + (p >> 3) * 2048 + v * 128 + (p & 7) * 16) = acc;
	}
	__syncthreads();
}
// the same from 128 basis images already in shared memory (img[i] = beta_i * z); no barrier inside
__device__ __forceinline__ void nlut_fill_from_images(uint8_t *tbl, const uint4 *img, uint32_t e) {
	const uint32_t p = e & 31, v = e >> 5;
	uint4 acc = u4_zero();
#pragma unroll
	for (uint32_t i = 0; i < 4; i++) {
		const uint32_t m = 0u - ((v >> i) & 1u);
		const uint4 w = img[4 * p + i];
		acc.x ^= w.x & m; acc.y ^= w.y & m; acc.z ^= w.z & m; acc.w ^= w.w & m;
	}
	*reinterpret_cast<uint4 *>(tbl + (p >> 3) * 2048 + v * 128 + (p & 7) * 16) = acc;
}
__device__ __forceinline__ uint4 nlut_ld(const uint8_t *tbl, uint32_t w, uint32_t n, uint32_t blk, uint32_t off) {
	// ((w >> 4n) & 15) * 128
	uint32_t r = n >= 2 ? ((w >> (4 * n - 7)) & 0x780u) : ((w << (7 - 4 * n)) & 0x780u);
	return *reinterpret_cast<const uint4 *>(tbl + blk * 2048u + r + off);
}
__device__ __forceinline__ uint4 nlut_apply(const uint8_t *tbl, const NLutLane &L, uint4 x) {
	uint32_t w[4] = {x.x, x.y, x.z, x.w};
	uint4 a0 = u4_zero(), a1 = u4_zero();
#pragma unroll
	for (uint32_t k = 0; k < 4; k++) {
		uint32_t v = __byte_perm(w[k], 0u, L.sel);
		uint32_t sw = ((v & 0x0F0F0F0Fu) << 4) | ((v >> 4) & 0x0F0F0F0Fu);
		v = L.nswap ? sw : v;
#pragma unroll
		for (uint32_t n = 0; n < 8; n += 2) {
			a0 ^= nlut_ld(tbl, v, n, k, L.off[n]);
			a1 ^= nlut_ld(tbl, v, n + 1, k, L.off[n + 1]);
		}
	}
	return a0 ^ a1;
}


// ------------------------------------------------------------------------------------------------
// Karatsuba-over-GF(2^64) variant of the byte-LUT engine ("K64").  With x = x0 + x1*X_6 and
// z = z0 + z1*X_6  (X_6^2 = X_6*X_5 + 1, pairwise_recursive_arithmetic.rs:12-30)
//     x*z = (z0 x0 + z1 x1) + (z1 x0 + (z0 + X_5 z1) x1) * X_6
// is a symmetric 2x2 matrix [[c0, c1], [c1, c2]] over GF(2^64) applied to (x0, x1), which needs only
// three products BY CONSTANTS and no post-multiplication:
//     P = c1 (x0 + x1),   lo = (c0 + c1) x0 + P,   hi = (c1 + c2) x1 + P.
// 3 x 8 byte-lookups of 8 bytes = 192 bytes of shared-memory reads per product instead of 16 x 16 = 256.
//
// Table layout (64 KiB): row b (256 B) = [ A_0..7[b] | B_0..7[b] | M_0..7[b] | M_0..7[b] ] with 8-byte
// entries; A = (.)*(c0+c1) on the bytes of x0, B = (.)*(c1+c2) on the bytes of x1, M = (.)*c1 on x0+x1.
// An LDS.64 is served per half-warp (16 lanes x 8 B = all 32 banks).  Lane (g, j) = ((lane>>3)&1,
// lane&7) walks byte positions k = s ^ j; lanes with g = 0 read A then B then M(copy 0), lanes with
// g = 1 read B then A then M(copy 1), so the 16 lanes of a half-warp always hit 16 distinct 8-byte
// bank pairs: conflict-free by construction.  Byte extraction and address formation are ONE PRMT:
// address = (byte << 8) | lane_offset with the offsets of four steps packed in a register.
struct K64Lane {
	uint32_t selP;      // PRMT selector: own operand's word (g) with bytes XOR-permuted by (j & 3)
	uint32_t selQ;      // same for the other operand
	uint32_t swap;      // swap the two words of each 64-bit half (j >> 2)
	uint32_t g;         // 0: own half is x0, 1: own half is x1
	uint32_t offP[2];   // byte s&3 of offP[s>>2] = g*64 + (s^j)*8       (own table set)
	uint32_t offQ[2];   //                         (1-g)*64 + (s^j)*8   (the other table set)
	uint32_t sbase;     // shared-window address of the table
};
__device__ __forceinline__ K64Lane k64_lane_init(const uint8_t *tbl) {
	K64Lane L;
	const uint32_t j = threadIdx.x & 7;
	L.g = (threadIdx.x >> 3) & 1u;
	const uint32_t perm = (j & 3u) * 0x1111u;
	L.selP = (L.g ? 0x7654u : 0x3210u) ^ perm;
	L.selQ = (L.g ? 0x3210u : 0x7654u) ^ perm;
	L.swap = (j >> 2) & 1u;
#pragma unroll
	for (uint32_t h = 0; h < 2; h++) {
		uint32_t p = 0, q = 0;
#pragma unroll
		for (uint32_t t = 0; t < 4; t++) {
			const uint32_t k = (4 * h + t) ^ j;
			p |= (L.g * 64u + k * 8u) << (8 * t);
			q |= ((1u - L.g) * 64u + k * 8u) << (8 * t);
		}
		L.offP[h] = p;
		L.offQ[h] = q;
	}
	L.sbase = (uint32_t)__cvta_generic_to_shared(tbl);
	return L;
}

// Build the tables of x -> x*z[t], t < R, side by side (tbl + t * LUT_BYTES).  stage: R * 192 uint2
// (1.5 KiB per table), laid out [bit i][set*8 + k] so that the entry loop reads and writes consecutive
// 8-byte words from consecutive lanes (conflict-free).  All threads of the CTA must call.
template <uint32_t R> __device__ __forceinline__ void k64_build_multi(uint8_t *tbl, uint2 *stage, const uint4 (&zs)[R]) {
	for (uint32_t e = threadIdx.x; e < R * 192; e += blockDim.x) {
		const uint32_t t = e / 192, r = e - t * 192, set = r >> 6, idx = r & 63;
		const uint4 z = zs[t];
		// c0 = z0, c1 = z1, c2 = z0 + X_5*z1;  A: c0+c1, B: c1+c2 = z0 + z1 + X_5*z1, M: c1
		const uint4 z1 = make_uint4(z.z, z.w, 0, 0);
		const uint4 az1 = mul_tower_gen(z1, 5);
		uint4 c = set == 0 ? make_uint4(z.x ^ z.z, z.y ^ z.w, 0, 0) : set == 1 ? make_uint4(z.x ^ z.z ^ az1.x, z.y ^ z.w ^ az1.y, 0, 0) : z1;
		uint4 im = basis_image(c, idx);  // idx < 64: stays inside the low GF(2^64) half
		stage[t * 192 + (idx & 7) * 24 + set * 8 + (idx >> 3)] = make_uint2(im.x, im.y);
	}
	__syncthreads();
	// Entry b of a column is the XOR of the images of the set bits of b.  Two passes: the 31 entries with one nibble
	// zero (b < 16 and b = 16 h) from the images, then every other entry as T[b & 15] ^ T[b & 0xF0] -- one XOR instead of
	// eight masked ones (the build is ~5 us of a CTA's issue slots the direct way; it is paid per table by
	// k_expand_outer and per round by the persistent sumcheck kernel).
	for (uint32_t e = threadIdx.x; e < R * 24 * 31; e += blockDim.x) {
		const uint32_t t = e / (24 * 31), r = e - t * (24 * 31);
		const uint32_t q = r / 24, col = r - q * 24;  // col = set*8 + k
		const uint32_t b = q < 16 ? q : (q - 15) << 4, nib = q < 16 ? q : q - 15, i0 = q < 16 ? 0 : 4;
		uint2 acc = make_uint2(0, 0);
#pragma unroll
		for (uint32_t i = 0; i < 4; i++) {
			const uint32_t m = 0u - ((nib >> i) & 1u);
			const uint2 w = stage[t * 192 + (i0 + i) * 24 + col];
			acc.x ^= w.x & m;
			acc.y ^= w.y & m;
		}
		uint2 *row = reinterpret_cast<uint2 *>(tbl + t * LUT_BYTES + b * 256u);
		row[col] = acc;
		if (col >= 16) row[col + 8] = acc;
	}
	__syncthreads();
	for (uint32_t e = threadIdx.x; e < R * 24 * 256; e += blockDim.x) {
		const uint32_t t = e / (24 * 256), r = e - t * (24 * 256);
		const uint32_t b = r / 24, col = r - b * 24;
		if (b < 16 || (b & 15u) == 0) continue;
		const uint8_t *base = tbl + t * LUT_BYTES;
		const uint2 lo = reinterpret_cast<const uint2 *>(base + (b & 15u) * 256u)[col], hi = reinterpret_cast<const uint2 *>(base + (b & 0xF0u) * 256u)[col];
		const uint2 acc = make_uint2(lo.x ^ hi.x, lo.y ^ hi.y);
		uint2 *row = reinterpret_cast<uint2 *>(tbl + t * LUT_BYTES + b * 256u);
		row[col] = acc;
		if (col >= 16) row[col + 8] = acc;
	}
	__syncthreads();
}
__device__ __forceinline__ void k64_build_mul(uint8_t *tbl, uint2 *stage, uint4 z) {
	const uint4 zs[1] = {z};
	k64_build_multi<1>(tbl, stage, zs);
}

template <uint32_t S, uint32_t IMM> __device__ __forceinline__ uint2 k64_ld(uint32_t sbase, uint32_t w, uint32_t off) {
	constexpr uint32_t t = S & 3u;
	constexpr uint32_t sel = ((0xCu + t) << 12) | ((0xCu + t) << 8) | (t << 4) | (4u + t);
	// (byte t of w) << 8 | (byte t of off); bytes 2-3 = replicated msb of the offset byte (< 128) = 0.
	// Inline PTX: __byte_perm masks the selector to 3 bits per nibble and drops the msb-replicate flag.
	uint32_t addr;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(addr) : "r"(w), "r"(off), "n"(sel));
	uint2 v;
	asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(addr + sbase), "n"(IMM));
	return v;
}
// 8 lookups of one 64-bit operand (w0, w1 already lane-permuted)
template <uint32_t IMM> __device__ __forceinline__ uint2 k64_set(uint32_t sbase, uint32_t w0, uint32_t w1, const uint32_t (&off)[2]) {
	uint2 a = k64_ld<0, IMM>(sbase, w0, off[0]), b = k64_ld<1, IMM>(sbase, w0, off[0]);
	uint2 c = k64_ld<2, IMM>(sbase, w0, off[0]), d = k64_ld<3, IMM>(sbase, w0, off[0]);
	uint2 e = k64_ld<4, IMM>(sbase, w1, off[1]), f = k64_ld<5, IMM>(sbase, w1, off[1]);
	uint2 g = k64_ld<6, IMM>(sbase, w1, off[1]), h = k64_ld<7, IMM>(sbase, w1, off[1]);
	return make_uint2((a.x ^ b.x ^ c.x) ^ (d.x ^ e.x ^ f.x) ^ (g.x ^ h.x), (a.y ^ b.y ^ c.y) ^ (d.y ^ e.y ^ f.y) ^ (g.y ^ h.y));
}
// y = x * z
__device__ __forceinline__ uint4 k64_apply(const K64Lane &L, uint4 x) {
	// word order by the lane's swap bit, then ONE PRMT per word picks own/other operand and permutes bytes
	const uint32_t u0 = L.swap ? x.y : x.x, u1 = L.swap ? x.x : x.y, v0 = L.swap ? x.w : x.z, v1 = L.swap ? x.z : x.w;
	const uint32_t a0 = __byte_perm(u0, v0, L.selP), a1 = __byte_perm(u1, v1, L.selP);
	const uint32_t b0 = __byte_perm(u0, v0, L.selQ), b1 = __byte_perm(u1, v1, L.selQ);
	const uint2 accP = k64_set<0>(L.sbase, a0, a1, L.offP);
	const uint2 accQ = k64_set<0>(L.sbase, b0, b1, L.offQ);
	const uint2 accM = k64_set<128>(L.sbase, a0 ^ b0, a1 ^ b1, L.offP);
	const uint2 lo = L.g ? accQ : accP, hi = L.g ? accP : accQ;  // A(x0), B(x1)
	return make_uint4(lo.x ^ accM.x, lo.y ^ accM.y, hi.x ^ accM.x, hi.y ^ accM.y);
}

}  // namespace b200
