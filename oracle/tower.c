/*
 * oracle/tower.c -- TEST INFRASTRUCTURE ONLY. See tower.h for the reference citations.
 */
#include "tower.h"

#include <string.h>

uint8_t TOWER_MUL8[256][256];
uint8_t TOWER_INV8[256];
uint8_t TOWER_ALPHA8[256];
static int g_init = 0;

static inline u128 lo_mask(int k) {
	/* mask of 2^k low bits, k in [0,7] */
	return k == 7 ? ~(u128)0 : (((u128)1 << (1u << k)) - 1);
}

/* pairwise_recursive_arithmetic.rs:48-62 -- multiply an element of T_k by X_{k-1}; identity in T_0 */
u128 tower_mul_alpha_slow(u128 a, int k) {
	if (k == 0) return a;
	int h = 1 << (k - 1);
	u128 m = lo_mask(k - 1);
	u128 a0 = a & m, a1 = (a >> h) & m;
	u128 z1 = tower_mul_alpha_slow(a1, k - 1);
	return a1 | ((a0 ^ z1) << h);
}

/* pairwise_recursive_arithmetic.rs:12-30 */
u128 tower_mul_slow(u128 a, u128 b, int k) {
	if (k == 0) return a & b & 1;
	int h = 1 << (k - 1);
	u128 m = lo_mask(k - 1);
	u128 a0 = a & m, a1 = (a >> h) & m;
	u128 b0 = b & m, b1 = (b >> h) & m;
	u128 z0 = tower_mul_slow(a0, b0, k - 1);
	u128 z2 = tower_mul_slow(a1, b1, k - 1);
	u128 z0z2 = z0 ^ z2;
	u128 z1 = tower_mul_slow(a0 ^ a1, b0 ^ b1, k - 1) ^ z0z2;
	u128 z2a = tower_mul_alpha_slow(z2, k - 1);
	return z0z2 | ((z1 ^ z2a) << h);
}

void tower_init(void) {
	if (g_init) return;
	for (int a = 0; a < 256; a++) {
		for (int b = a; b < 256; b++) {
			uint8_t p = (uint8_t)tower_mul_slow((u128)a, (u128)b, 3);
			TOWER_MUL8[a][b] = p;
			TOWER_MUL8[b][a] = p;
		}
		TOWER_ALPHA8[a] = (uint8_t)tower_mul_alpha_slow((u128)a, 3);
	}
	TOWER_INV8[0] = 0;
	for (int a = 1; a < 256; a++)
		for (int b = 1; b < 256; b++)
			if (TOWER_MUL8[a][b] == 1) { TOWER_INV8[a] = (uint8_t)b; break; }
	g_init = 1;
}

u128 tower_mul_alpha(u128 a, int k) {
	if (k < 3) return tower_mul_alpha_slow(a, k);
	if (k == 3) return TOWER_ALPHA8[(uint8_t)a];
	int h = 1 << (k - 1);
	u128 m = lo_mask(k - 1);
	u128 a0 = a & m, a1 = (a >> h) & m;
	return a1 | ((a0 ^ tower_mul_alpha(a1, k - 1)) << h);
}

u128 tower_mul(u128 a, u128 b, int k) {
	if (k < 3) return tower_mul_slow(a, b, k);
	if (k == 3) return TOWER_MUL8[(uint8_t)a][(uint8_t)b];
	int h = 1 << (k - 1);
	u128 m = lo_mask(k - 1);
	u128 a0 = a & m, a1 = (a >> h) & m;
	u128 b0 = b & m, b1 = (b >> h) & m;
	u128 z0 = tower_mul(a0, b0, k - 1);
	u128 z2 = tower_mul(a1, b1, k - 1);
	u128 z0z2 = z0 ^ z2;
	u128 z1 = tower_mul(a0 ^ a1, b0 ^ b1, k - 1) ^ z0z2;
	return z0z2 | ((z1 ^ tower_mul_alpha(z2, k - 1)) << h);
}

/* pairwise_recursive_arithmetic.rs:33-45 */
u128 tower_square(u128 a, int k) {
	if (k <= 3) return tower_mul(a, a, k);
	int h = 1 << (k - 1);
	u128 m = lo_mask(k - 1);
	u128 a0 = a & m, a1 = (a >> h) & m;
	u128 z0 = tower_square(a0, k - 1), z2 = tower_square(a1, k - 1);
	return (z0 ^ z2) | (tower_mul_alpha(z2, k - 1) << h);
}

/* pairwise_recursive_arithmetic.rs:65-81 */
u128 tower_invert(u128 a, int k) {
	if (k == 0) return a & 1;
	if (k == 3) return TOWER_INV8[(uint8_t)a];
	int h = 1 << (k - 1);
	u128 m = lo_mask(k - 1);
	u128 a0 = a & m, a1 = (a >> h) & m;
	u128 a0z1 = a0 ^ tower_mul_alpha(a1, k - 1);
	u128 delta = tower_mul(a0, a0z1, k - 1) ^ tower_square(a1, k - 1);
	u128 dinv = tower_invert(delta, k - 1);
	return tower_mul(dinv, a0z1, k - 1) | (tower_mul(dinv, a1, k - 1) << h);
}

/* binary_field.rs:363-414 */
u128 tower_mul_subfield(u128 a, u128 s, int k) {
	if (k == 7) return tower_mul(a, s, 7);
	if (k == 0) return (s & 1) ? a : 0;
	int w = 1 << k;
	u128 m = lo_mask(k);
	u128 r = 0;
	for (int sh = 0; sh < 128; sh += w) r |= tower_mul((a >> sh) & m, s & m, k) << sh;
	return r;
}
