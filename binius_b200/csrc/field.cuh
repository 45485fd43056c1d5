// field.cuh -- device-side binary tower field arithmetic (general, per-lane).
//
// Semantics: reference crates/field/src/arch/portable/pairwise_recursive_arithmetic.rs:12-62
// (Karatsuba tower step, mul_alpha) with the 8-bit level served from shared-memory tables
// (the reference's PairwiseTableStrategy idea, pairwise_table_arithmetic.rs:57-120; our tables are
// generated at context creation from the bit-level recursion in host_field.hpp).
//
// A B128 element is a uint4 {x = bits 0..31, y, z, w = bits 96..127}.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

// shared-memory resident tables: 64 KiB product table of T_3 (B8) + 256 B "times X_2" table
struct FieldTables {
	const uint8_t *mul8;    // [256][256]
	const uint8_t *alpha8;  // [256]
};
constexpr uint32_t FIELD_TABLE_BYTES = 65536 + 256;

// cooperative copy of the global tables into shared memory (all threads of the CTA must call)
__device__ __forceinline__ FieldTables load_field_tables(uint8_t *smem, const uint8_t *g_tables) {
	const uint4 *src = reinterpret_cast<const uint4 *>(g_tables);
	uint4 *dst = reinterpret_cast<uint4 *>(smem);
	for (uint32_t i = threadIdx.x; i < FIELD_TABLE_BYTES / 16; i += blockDim.x) dst[i] = __ldg(src + i);
	__syncthreads();
	FieldTables t;
	t.mul8 = smem;
	t.alpha8 = smem + 65536;
	return t;
}

__device__ __forceinline__ uint4 operator^(uint4 a, uint4 b) { return make_uint4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w); }
__device__ __forceinline__ uint4 &operator^=(uint4 &a, uint4 b) {
	a.x ^= b.x; a.y ^= b.y; a.z ^= b.z; a.w ^= b.w;
	return a;
}
__device__ __forceinline__ uint2 operator^(uint2 a, uint2 b) { return make_uint2(a.x ^ b.x, a.y ^ b.y); }
__device__ __forceinline__ bool is_zero(uint4 a) { return (a.x | a.y | a.z | a.w) == 0; }
__device__ __forceinline__ uint4 u4_zero() { return make_uint4(0, 0, 0, 0); }
__device__ __forceinline__ uint4 u4_one() { return make_uint4(1, 0, 0, 0); }

// Bank behaviour of the 64 KiB table: the FIRST operand picks the 256-byte row (64 words = every bank twice), the second
// the byte inside it.  When one operand is the same in every lane of the warp (a challenge, an evaluation point, a
// sub-field scalar) it must be the FIRST one: all lanes then read one row (at most 2-way conflicts); as the second
// operand it would put all 32 lanes on ONE bank (32-way conflict, measured 5x on a fold).  Two per-lane operands see
// ~3.5-way conflicts either way.
__device__ __forceinline__ uint32_t f_mul8(const FieldTables &T, uint32_t a, uint32_t b) { return T.mul8[(a << 8) | b]; }

__device__ __forceinline__ uint32_t f_alpha16(const FieldTables &T, uint32_t a) {
	uint32_t a0 = a & 0xff, a1 = a >> 8;
	return a1 | ((a0 ^ T.alpha8[a1]) << 8);
}
__device__ __forceinline__ uint32_t f_mul16(const FieldTables &T, uint32_t a, uint32_t b) {
	uint32_t a0 = a & 0xff, a1 = a >> 8, b0 = b & 0xff, b1 = b >> 8;
	uint32_t p0 = f_mul8(T, a0, b0), p2 = f_mul8(T, a1, b1);
	uint32_t pm = f_mul8(T, a0 ^ a1, b0 ^ b1);
	return (p0 ^ p2) | ((pm ^ p0 ^ p2 ^ T.alpha8[p2]) << 8);
}
__device__ __forceinline__ uint32_t f_alpha32(const FieldTables &T, uint32_t a) {
	uint32_t a0 = a & 0xffff, a1 = a >> 16;
	return a1 | ((a0 ^ f_alpha16(T, a1)) << 16);
}
__device__ __forceinline__ uint32_t f_mul32(const FieldTables &T, uint32_t a, uint32_t b) {
	uint32_t a0 = a & 0xffff, a1 = a >> 16, b0 = b & 0xffff, b1 = b >> 16;
	uint32_t p0 = f_mul16(T, a0, b0), p2 = f_mul16(T, a1, b1);
	uint32_t pm = f_mul16(T, a0 ^ a1, b0 ^ b1);
	return (p0 ^ p2) | ((pm ^ p0 ^ p2 ^ f_alpha16(T, p2)) << 16);
}
__device__ __forceinline__ uint2 f_alpha64(const FieldTables &T, uint2 a) { return make_uint2(a.y, a.x ^ f_alpha32(T, a.y)); }
__device__ __forceinline__ uint2 f_mul64(const FieldTables &T, uint2 a, uint2 b) {
	uint32_t p0 = f_mul32(T, a.x, b.x), p2 = f_mul32(T, a.y, b.y);
	uint32_t pm = f_mul32(T, a.x ^ a.y, b.x ^ b.y);
	return make_uint2(p0 ^ p2, pm ^ p0 ^ p2 ^ f_alpha32(T, p2));
}
__device__ __noinline__ uint4 f_mul128(const FieldTables &T, uint4 a, uint4 b) {
	uint2 a0 = make_uint2(a.x, a.y), a1 = make_uint2(a.z, a.w);
	uint2 b0 = make_uint2(b.x, b.y), b1 = make_uint2(b.z, b.w);
	uint2 p0 = f_mul64(T, a0, b0), p2 = f_mul64(T, a1, b1);
	uint2 pm = f_mul64(T, a0 ^ a1, b0 ^ b1);
	uint2 lo = p0 ^ p2;
	uint2 hi = pm ^ lo ^ f_alpha64(T, p2);
	return make_uint4(lo.x, lo.y, hi.x, hi.y);
}

// B128 x subfield element of tower level `lvl` (lvl in {0,3,4,5,6,7}), limb-wise
// (reference crates/field/src/binary_field.rs:363-414)
__device__ __forceinline__ uint32_t f_mul8x4(const FieldTables &T, uint32_t a, uint32_t s) {
	return f_mul8(T, s, a & 0xff) | (f_mul8(T, s, (a >> 8) & 0xff) << 8) | (f_mul8(T, s, (a >> 16) & 0xff) << 16) | (f_mul8(T, s, a >> 24) << 24);
}
__device__ __forceinline__ uint32_t f_mul16x2(const FieldTables &T, uint32_t a, uint32_t s) {
	return f_mul16(T, s, a & 0xffff) | (f_mul16(T, s, a >> 16) << 16);
}
// B128 x B8 with a WARP-UNIFORM B128 operand and a per-lane B8 scalar (the eq-indicator tables of the univariate-skip
// round: entry e of a table is eq * e): the uniform bytes pick the table row, so the lanes of a warp read one 256-byte
// row (<= 2-way conflicts) instead of 32 rows in the same bank.
__device__ __forceinline__ uint32_t f_mul8x4_uniform(const FieldTables &T, uint32_t a, uint32_t s) {
	return f_mul8(T, a & 0xff, s) | (f_mul8(T, (a >> 8) & 0xff, s) << 8) | (f_mul8(T, (a >> 16) & 0xff, s) << 16) | (f_mul8(T, a >> 24, s) << 24);
}
__device__ __forceinline__ uint4 f_mul128_b8_uniform(const FieldTables &T, uint4 a, uint32_t s) {
	return make_uint4(f_mul8x4_uniform(T, a.x, s), f_mul8x4_uniform(T, a.y, s), f_mul8x4_uniform(T, a.z, s), f_mul8x4_uniform(T, a.w, s));
}
// (the scalar s, often warp-uniform, goes first: see f_mul8)
__device__ __forceinline__ uint4 f_mul128_sub(const FieldTables &T, uint4 a, uint4 s, uint32_t lvl) {
	switch (lvl) {
	case 0: {
		uint32_t m = 0u - (s.x & 1u);
		return make_uint4(a.x & m, a.y & m, a.z & m, a.w & m);
	}
	case 3: {
		uint32_t c = s.x & 0xff;
		return make_uint4(f_mul8x4(T, a.x, c), f_mul8x4(T, a.y, c), f_mul8x4(T, a.z, c), f_mul8x4(T, a.w, c));
	}
	case 4: {
		uint32_t c = s.x & 0xffff;
		return make_uint4(f_mul16x2(T, a.x, c), f_mul16x2(T, a.y, c), f_mul16x2(T, a.z, c), f_mul16x2(T, a.w, c));
	}
	case 5: return make_uint4(f_mul32(T, s.x, a.x), f_mul32(T, s.x, a.y), f_mul32(T, s.x, a.z), f_mul32(T, s.x, a.w));
	case 6: {
		uint2 c = make_uint2(s.x, s.y);
		uint2 l = f_mul64(T, c, make_uint2(a.x, a.y)), h = f_mul64(T, c, make_uint2(a.z, a.w));
		return make_uint4(l.x, l.y, h.x, h.y);
	}
	default: return f_mul128(T, s, a);
	}
}

// limb j (2^lvl bits, low limb first) of a B128, zero-extended  (binary_field.rs:600-607, 628-657)
__device__ __forceinline__ uint4 f_limb(uint4 a, uint32_t lvl, uint32_t j) {
	uint32_t w[4] = {a.x, a.y, a.z, a.w};
	switch (lvl) {
	case 0: return make_uint4((w[j >> 5] >> (j & 31)) & 1u, 0, 0, 0);
	case 3: return make_uint4((w[j >> 2] >> ((j & 3) * 8)) & 0xffu, 0, 0, 0);
	case 4: return make_uint4((w[j >> 1] >> ((j & 1) * 16)) & 0xffffu, 0, 0, 0);
	case 5: return make_uint4(w[j], 0, 0, 0);
	case 6: return make_uint4(w[2 * j], w[2 * j + 1], 0, 0);
	default: return a;
	}
}

// warp + block XOR reduction of a uint4; result valid in thread 0. `red` = 32 uint4 of smem.
__device__ __forceinline__ uint4 warp_xor(uint4 v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		v.x ^= __shfl_xor_sync(0xffffffffu, v.x, o);
		v.y ^= __shfl_xor_sync(0xffffffffu, v.y, o);
		v.z ^= __shfl_xor_sync(0xffffffffu, v.z, o);
		v.w ^= __shfl_xor_sync(0xffffffffu, v.w, o);
	}
	return v;
}
__device__ __forceinline__ uint4 block_xor(uint4 v, uint4 *red) {
	v = warp_xor(v);
	uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	__syncthreads();
	if (lane == 0) red[warp] = v;
	__syncthreads();
	if (warp == 0) {
		v = lane < nw ? red[lane] : u4_zero();
		v = warp_xor(v);
	}
	return v;
}
__device__ __forceinline__ void atomic_xor_u4(uint4 *dst, uint4 v) {
	uint32_t *d = reinterpret_cast<uint32_t *>(dst);
	if (v.x) atomicXor(d + 0, v.x);
	if (v.y) atomicXor(d + 1, v.y);
	if (v.z) atomicXor(d + 2, v.z);
	if (v.w) atomicXor(d + 3, v.w);
}

}  // namespace b200
