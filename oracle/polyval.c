/*
 * oracle/polyval.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of BinaryField128bPolyval, the "fast" field of the reference's GKR grand-product prover:
 *   crates/field/src/arch/portable/packed_polyval_128.rs:88-160   montgomery_multiply (a * b * X^-128 in
 *       GF(2)[X] / (X^128 + X^127 + X^126 + X^121 + 1)), bmul64 (carry-less 64 x 64 -> low 64 bits with holes), rev64
 *   crates/field/src/polyval.rs:262 ONE, :308-311 to_montgomery, :516-788 tower <-> POLYVAL basis change tables
 *   crates/core/src/protocols/gkr_gpa/gkr_gpa.rs:40-90   GrandProductWitness::new (layer k = lo half * hi half)
 * Pinned by the reference's own KATs (polyval.rs:1113-1127, tests/golden/field_kat.json) in tests/test_oracle_field.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "tower.h"

typedef u128 __attribute__((aligned(8))) u128u;

static uint64_t bmul64(uint64_t x, uint64_t y) {
	uint64_t x0 = x & 0x1111111111111111ull, x1 = x & 0x2222222222222222ull, x2 = x & 0x4444444444444444ull, x3 = x & 0x8888888888888888ull;
	uint64_t y0 = y & 0x1111111111111111ull, y1 = y & 0x2222222222222222ull, y2 = y & 0x4444444444444444ull, y3 = y & 0x8888888888888888ull;
	uint64_t z0 = (x0 * y0) ^ (x1 * y3) ^ (x2 * y2) ^ (x3 * y1);
	uint64_t z1 = (x0 * y1) ^ (x1 * y0) ^ (x2 * y3) ^ (x3 * y2);
	uint64_t z2 = (x0 * y2) ^ (x1 * y1) ^ (x2 * y0) ^ (x3 * y3);
	uint64_t z3 = (x0 * y3) ^ (x1 * y2) ^ (x2 * y1) ^ (x3 * y0);
	return (z0 & 0x1111111111111111ull) | (z1 & 0x2222222222222222ull) | (z2 & 0x4444444444444444ull) | (z3 & 0x8888888888888888ull);
}
static uint64_t rev64(uint64_t x) {
	x = ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);
	x = ((x & 0x3333333333333333ull) << 2) | ((x >> 2) & 0x3333333333333333ull);
	x = ((x & 0x0f0f0f0f0f0f0f0full) << 4) | ((x >> 4) & 0x0f0f0f0f0f0f0f0full);
	x = ((x & 0x00ff00ff00ff00ffull) << 8) | ((x >> 8) & 0x00ff00ff00ff00ffull);
	x = ((x & 0x0000ffff0000ffffull) << 16) | ((x >> 16) & 0x0000ffff0000ffffull);
	return (x << 32) | (x >> 32);
}

/* packed_polyval_128.rs:88-129 */
u128 polyval_mul(u128 a, u128 b) {
	uint64_t h0 = (uint64_t)a, h1 = (uint64_t)(a >> 64), h0r = rev64(h0), h1r = rev64(h1), h2 = h0 ^ h1, h2r = h0r ^ h1r;
	uint64_t y0 = (uint64_t)b, y1 = (uint64_t)(b >> 64), y0r = rev64(y0), y1r = rev64(y1), y2 = y0 ^ y1, y2r = y0r ^ y1r;
	uint64_t z0 = bmul64(y0, h0), z1 = bmul64(y1, h1), z2 = bmul64(y2, h2);
	uint64_t z0h = bmul64(y0r, h0r), z1h = bmul64(y1r, h1r), z2h = bmul64(y2r, h2r);
	z2 ^= z0 ^ z1;
	z2h ^= z0h ^ z1h;
	z0h = rev64(z0h) >> 1;
	z1h = rev64(z1h) >> 1;
	z2h = rev64(z2h) >> 1;
	uint64_t v0 = z0, v1 = z0h ^ z2, v2 = z1 ^ z2h, v3 = z1h;
	v2 ^= v0 ^ (v0 >> 1) ^ (v0 >> 2) ^ (v0 >> 7);
	v1 ^= (v0 << 63) ^ (v0 << 62) ^ (v0 << 57);
	v3 ^= v1 ^ (v1 >> 1) ^ (v1 >> 2) ^ (v1 >> 7);
	v2 ^= (v1 << 63) ^ (v1 << 62) ^ (v1 << 57);
	return (u128)v2 | ((u128)v3 << 64);
}

void orc_polyval_mul(const u128u *a, const u128u *b, u128u *out) { *out = polyval_mul(*a, *b); }
void orc_polyval_mul_vec(const u128u *a, const u128u *b, u128u *out, uint64_t n) {
	for (uint64_t i = 0; i < n; i++) out[i] = polyval_mul(a[i], b[i]);
}
/* FieldLinearTransformation::transform: out = XOR of images[k] over the set bits k of x */
void orc_linear_map(const u128u *images /* 128 */, const u128u *x, u128u *out, uint64_t n) {
	for (uint64_t i = 0; i < n; i++) {
		u128 v = x[i], acc = 0;
		for (int k = 0; k < 128; k++)
			if ((v >> k) & 1) acc ^= images[k];
		out[i] = acc;
	}
}
/* GrandProductWitness::new (gkr_gpa.rs:40-90) over full layers of POLYVAL elements: layers[0] = input (2^n_vars),
 * layer k+1 [i] = layer k [i] * layer k [i + len/2]; all layers concatenated into `out` (2^(n_vars+1) - 1 elements). */
void orc_polyval_gpa_layers(const u128u *input, uint32_t n_vars, u128u *out) {
	uint64_t len = (uint64_t)1 << n_vars;
	memcpy(out, input, sizeof(u128) * len);
	const u128u *prev = out;
	u128u *cur = out + len;
	while (len > 1) {
		uint64_t half = len / 2;
		for (uint64_t i = 0; i < half; i++) cur[i] = polyval_mul(prev[i], prev[half + i]);
		prev = cur;
		cur += half;
		len = half;
	}
}
/* eq-ind round values of the GPA layer sumcheck in POLYVAL arithmetic (composition = product of the two half-layer
 * multilinears, hal/src/sumcheck_round_calculation.rs:222-297 + core/.../prove/eq_ind.rs:646-731):
 *   at 1: sum_i E[i] * A.hi[i] * B.hi[i];   at infinity: sum_i E[i] * (A.hi - A.lo)[i] * (B.hi - B.lo)[i]   */
void orc_polyval_gpa_round_evals(const u128u *a, const u128u *b, const u128u *eq, uint32_t n_vars, u128u *out2) {
	uint64_t half = (uint64_t)1 << (n_vars - 1);
	u128 y1 = 0, yinf = 0;
	for (uint64_t i = 0; i < half; i++) {
		y1 ^= polyval_mul(eq[i], polyval_mul(a[half + i], b[half + i]));
		yinf ^= polyval_mul(eq[i], polyval_mul(a[half + i] ^ a[i], b[half + i] ^ b[i]));
	}
	out2[0] = y1;
	out2[1] = yinf;
}
