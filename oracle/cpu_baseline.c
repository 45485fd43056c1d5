/*
 * oracle/cpu_baseline.c -- TEST / MEASUREMENT INFRASTRUCTURE ONLY (bench.py CPU arm).
 *
 * "Port" of the reference's fastest x86 path for the fold-high loop, restated in C because the
 * reference (Rust nightly) cannot be built in this image:
 *   - loop:      fold_left_lerp_inplace, crates/math/src/fold.rs:648-696 (e0 += (e1 - e0) * z), chunked
 *                over threads the way FastCpuLayer::extrapolate_line does with rayon
 *                (crates/fast_compute/src/layer.rs:515-550)
 *   - multiply:  the GFNI strategy of crates/field/src/arch/x86_64/gfni/{gfni_arithmetics.rs:22-103,
 *                aes_isomorphic.rs:38-98}: GF2P8AFFINEQB tower->AES basis change on every byte, packed
 *                Karatsuba tower levels on 512-bit registers (simd/simd_arithmetic.rs:180-220) with
 *                GF2P8MULB as the 8-bit base multiply, GF2P8AFFINEQB back.
 * Labelled everywhere as "C restatement of the reference GFNI path -- not the Rust binary".
 * Falls back to the scalar oracle loop (threaded) when the host CPU lacks AVX-512BW + GFNI.
 * Validated against the scalar oracle in tests/test_oracle_ops.py::test_cpu_baseline_matches_oracle.
 */
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "tower.h"

typedef u128 __attribute__((aligned(8))) u128u;

/* byte-wise tower <-> AES isomorphism (crates/field/src/aes_field.rs:113-141): images of the basis */
static const uint8_t TOWER_TO_AES[8] = {0x01, 0xbc, 0xb0, 0xec, 0xd3, 0x8d, 0x2e, 0x58};
static const uint8_t AES_TO_TOWER[8] = {0x01, 0x3c, 0x8c, 0x8a, 0x59, 0x7a, 0x53, 0x27};

/* GF2P8AFFINEQB matrix: result bit i = parity(A.byte[7-i] & x); row_i bit j = bit i of image(e_j) */
static uint64_t affine_matrix(const uint8_t img[8]) {
	uint64_t m = 0;
	for (int i = 0; i < 8; i++) {
		uint8_t row = 0;
		for (int j = 0; j < 8; j++) row |= (uint8_t)(((img[j] >> i) & 1) << j);
		m |= (uint64_t)row << (8 * (7 - i));
	}
	return m;
}

int cpu_has_gfni512(void) {
	return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("gfni");
}

#define TGT __attribute__((target("avx512f,avx512bw,avx512vl,gfni")))

/* swap the two halves (h bytes each) of every 2h-byte element */
TGT static inline __m512i swap_halves(__m512i x, int k) {
	switch (k) {
	case 4: {
		const __m512i idx = _mm512_set4_epi32(0x0e0f0c0d, 0x0a0b0809, 0x06070405, 0x02030001);
		return _mm512_shuffle_epi8(x, idx);
	}
	case 5: return _mm512_rol_epi32(x, 16);
	case 6: return _mm512_shuffle_epi32(x, (_MM_PERM_ENUM)0xB1);
	default: return _mm512_shuffle_epi32(x, (_MM_PERM_ENUM)0x4E);
	}
}
/* byte mask selecting the high half of every 2^k-bit element */
static inline __mmask64 hi_mask(int k) {
	switch (k) {
	case 4: return 0xAAAAAAAAAAAAAAAAull;
	case 5: return 0xCCCCCCCCCCCCCCCCull;
	case 6: return 0xF0F0F0F0F0F0F0F0ull;
	default: return 0xFF00FF00FF00FF00ull;
	}
}
/* multiply every 2^k-bit element (AES-tower representation) by X_{k-1}; one function per level so
 * that everything inlines (the reference monomorphises per packed type) */
TGT static inline __m512i aes_alpha3(__m512i x) { return _mm512_gf2p8mul_epi8(x, _mm512_set1_epi8((char)0xD3)); }
TGT static inline __m512i aes_mul3(__m512i a, __m512i b) { return _mm512_gf2p8mul_epi8(a, b); }
#define DEF_LEVEL(K, KM1)                                                                                  \
	TGT static inline __m512i aes_alpha##K(__m512i x) {                                                    \
		__m512i sw = swap_halves(x, K);                                                                    \
		__m512i al = aes_alpha##KM1(x);                                                                    \
		return _mm512_xor_si512(sw, _mm512_maskz_mov_epi8(hi_mask(K), al));                                \
	}                                                                                                      \
	/* packed Karatsuba tower step (simd_arithmetic.rs:180-220) */                                         \
	TGT static inline __m512i aes_mul##K(__m512i a, __m512i b) {                                           \
		__m512i as = _mm512_xor_si512(a, swap_halves(a, K));                                               \
		__m512i bs = _mm512_xor_si512(b, swap_halves(b, K));                                               \
		__m512i z02 = aes_mul##KM1(a, b);                                                                  \
		__m512i z1f = aes_mul##KM1(as, bs);                                                                \
		__m512i z02s = _mm512_xor_si512(z02, swap_halves(z02, K));                                         \
		__m512i hi = _mm512_xor_si512(_mm512_xor_si512(z1f, z02s), aes_alpha##KM1(z02));                   \
		return _mm512_mask_blend_epi8(hi_mask(K), z02s, hi);                                               \
	}
DEF_LEVEL(4, 3)
DEF_LEVEL(5, 4)
DEF_LEVEL(6, 5)
DEF_LEVEL(7, 6)

TGT static void fold_gfni(u128u *e0, const u128u *e1, uint64_t n, u128 z) {
	const __m512i t2a = _mm512_set1_epi64((long long)affine_matrix(TOWER_TO_AES));
	const __m512i a2t = _mm512_set1_epi64((long long)affine_matrix(AES_TO_TOWER));
	__m512i zv = _mm512_broadcast_i32x4(_mm_loadu_si128((const __m128i *)&z));
	zv = _mm512_gf2p8affine_epi64_epi8(zv, t2a, 0);
	uint64_t i = 0;
	for (; i + 4 <= n; i += 4) {
		__m512i a = _mm512_loadu_si512((const void *)(e0 + i));
		__m512i b = _mm512_loadu_si512((const void *)(e1 + i));
		__m512i d = _mm512_gf2p8affine_epi64_epi8(_mm512_xor_si512(a, b), t2a, 0);
		__m512i p = _mm512_gf2p8affine_epi64_epi8(aes_mul7(d, zv), a2t, 0);
		_mm512_storeu_si512((void *)(e0 + i), _mm512_xor_si512(a, p));
	}
	for (; i < n; i++) e0[i] ^= b128_mul(e1[i] ^ e0[i], z);
}

static void fold_scalar(u128u *e0, const u128u *e1, uint64_t n, u128 z) {
	for (uint64_t i = 0; i < n; i++) e0[i] ^= b128_mul(e1[i] ^ e0[i], z);
}

typedef struct {
	u128u *e0;
	const u128u *e1;
	uint64_t n;
	u128 z;
	int gfni;
} fold_job;

static void *fold_worker(void *p) {
	fold_job *j = (fold_job *)p;
	if (j->gfni) fold_gfni(j->e0, j->e1, j->n, j->z);
	else fold_scalar(j->e0, j->e1, j->n, j->z);
	return NULL;
}

/* e0[i] += (e1[i] - e0[i]) * z over n elements with n_threads threads; use_gfni = 0 forces scalar */
int cpu_fold(u128u *e0, const u128u *e1, uint64_t n, const u128u *z, int n_threads, int use_gfni) {
	tower_init();
	int gfni = use_gfni && cpu_has_gfni512();
	if (n_threads < 1) n_threads = 1;
	if (n_threads == 1) {
		fold_job j = {e0, e1, n, *z, gfni};
		fold_worker(&j);
		return gfni;
	}
	pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
	fold_job *jobs = malloc(sizeof(fold_job) * n_threads);
	uint64_t per = ((n + n_threads - 1) / n_threads + 3) & ~3ull;
	int started = 0;
	for (int t = 0; t < n_threads; t++) {
		uint64_t s = (uint64_t)t * per;
		if (s >= n) break;
		uint64_t cnt = n - s < per ? n - s : per;
		jobs[t] = (fold_job){e0 + s, e1 + s, cnt, *z, gfni};
		pthread_create(&th[t], NULL, fold_worker, &jobs[t]);
		started++;
	}
	for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
	free(th);
	free(jobs);
	return gfni;
}

/* timed loop for bench.py: folds a 2^log_n-coefficient multilinear `reps` times in place, returns seconds */
double cpu_fold_bench(uint32_t log_n, int reps, int n_threads, int use_gfni, int *used_gfni) {
	uint64_t half = (uint64_t)1 << (log_n - 1);
	u128u *buf = aligned_alloc(64, sizeof(u128) * 2 * half);
	uint64_t s = 0x1234567;
	for (uint64_t i = 0; i < 2 * half; i++) {
		s = s * 6364136223846793005ull + 1442695040888963407ull;
		buf[i] = ((u128)s << 64) | (s * 0x9E3779B97F4A7C15ull);
	}
	u128 z = ((u128)0x2E895399AF449ACEull << 64) | 0x499596F6E5FCCAFAull;
	cpu_fold(buf, buf + half, half, (const u128u *)&z, n_threads, use_gfni); /* warm-up, page-in */
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	int g = 0;
	for (int r = 0; r < reps; r++) g = cpu_fold(buf, buf + half, half, (const u128u *)&z, n_threads, use_gfni);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (used_gfni) *used_gfni = g;
	free(buf);
	return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* timed loop for bench.py "sumcheck_chain": folds a 2^log_n-coefficient multilinear all the way down
 * (log_n rounds of fold-high, sizes 2^log_n, 2^(log_n-1), ..., 2; the reference prover's per-multilinear
 * work over one sumcheck, sumcheck_folding.rs:223-237), `reps` times; returns seconds.  Rounds below
 * 2^14 coefficients run on one thread (thread start-up would dominate). */
double cpu_fold_chain_bench(uint32_t log_n, int reps, int n_threads, int use_gfni) {
	uint64_t n = (uint64_t)1 << log_n;
	u128u *src = aligned_alloc(64, sizeof(u128) * n), *buf = aligned_alloc(64, sizeof(u128) * n);
	uint64_t s = 0x1234567;
	for (uint64_t i = 0; i < n; i++) {
		s = s * 6364136223846793005ull + 1442695040888963407ull;
		src[i] = ((u128)s << 64) | (s * 0x9E3779B97F4A7C15ull);
	}
	u128 z = ((u128)0x2E895399AF449ACEull << 64) | 0x499596F6E5FCCAFAull;
	double total = 0;
	for (int r = -1; r < reps; r++) { /* r = -1: warm-up */
		memcpy(buf, src, sizeof(u128) * n);
		struct timespec t0, t1;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		for (uint32_t v = log_n; v >= 1; v--) {
			uint64_t half = (uint64_t)1 << (v - 1);
			cpu_fold(buf, buf + half, half, (const u128u *)&z, half >= (1u << 13) ? n_threads : 1, use_gfni);
			z = z * 3 + 1; /* a different challenge every round */
		}
		clock_gettime(CLOCK_MONOTONIC, &t1);
		if (r >= 0) total += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
	}
	free(src);
	free(buf);
	return total;
}
