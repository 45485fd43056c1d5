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

/* =================================================================================================
 * CPU arm of the additive NTT over B32 (BASELINE config #2 and the RS-encode phase of the keccak replay).
 * "Port" of the reference's multithreaded NTT, crates/ntt/src/multithreaded.rs:100-228:
 *   phase 1  the top `par_rounds` layers as strided column transforms, one stride of packed columns per thread
 *   phase 2  one single-threaded transform (single_threaded.rs:134-246) per row chunk, coset = the row index
 * on PackedBinaryField16x32b = one __m512i, twiddle broadcast per block (precomputed per-layer tables,
 * twiddle.rs:324-353), multiply = the GFNI strategy above at tower level 5.  Requires log_x >= 4 (every run of
 * butterflies sharing a twiddle is at least one 512-bit register long), which covers S1 and the RS-encode shapes.
 * Checked against the scalar oracle NTT in tests/test_oracle_ops.py.
 * ================================================================================================= */

typedef struct {
	uint32_t *data;
	uint32_t log_x, log_y, skip;
	const uint32_t *const *tw; /* tw[I][j]: twiddle of global layer I, block j (coset 0), I < log_y */
	uint32_t par_rounds, log_rows_elems; /* phase 1: rows of 2^log_row elements */
	uint64_t begin, end;                 /* phase 1: element-column range [begin, end) in units of 16 scalars; phase 2: chunk range */
	int phase, gfni;
	uint64_t t2a, a2t;
} ntt_job;

TGT static void ntt_butterfly_run_gfni(uint32_t *u, uint32_t *v, uint64_t n16, uint32_t t, uint64_t t2a, uint64_t a2t) {
	const __m512i mt = _mm512_set1_epi64((long long)t2a), mb = _mm512_set1_epi64((long long)a2t);
	const __m512i tv = _mm512_gf2p8affine_epi64_epi8(_mm512_set1_epi32((int)t), mt, 0);
	for (uint64_t i = 0; i < n16; i++) {
		__m512i a = _mm512_loadu_si512((const void *)(u + 16 * i));
		__m512i b = _mm512_loadu_si512((const void *)(v + 16 * i));
		__m512i p = _mm512_gf2p8affine_epi64_epi8(aes_mul5(_mm512_gf2p8affine_epi64_epi8(b, mt, 0), tv), mb, 0);
		a = _mm512_xor_si512(a, p);
		b = _mm512_xor_si512(b, a);
		_mm512_storeu_si512((void *)(u + 16 * i), a);
		_mm512_storeu_si512((void *)(v + 16 * i), b);
	}
}
static void ntt_butterfly_run_scalar(uint32_t *u, uint32_t *v, uint64_t n, uint32_t t) {
	for (uint64_t i = 0; i < n; i++) {
		u[i] ^= (uint32_t)tower_mul(v[i], t, 5);
		v[i] ^= u[i];
	}
}

static void *ntt_worker(void *p) {
	ntt_job *J = (ntt_job *)p;
	const uint32_t lx = J->log_x, ly = J->log_y;
	if (J->phase == 1) {
		/* rows = the top par_rounds bits of y; a row holds 2^(lx + ly - par_rounds) scalars; this thread owns the
		 * 16-scalar columns [begin, end) of every row */
		const uint32_t pr = J->par_rounds, log_row = lx + ly - pr;
		for (int i = (int)pr - 1 - (int)J->skip; i >= 0; i--) {
			const uint32_t I = ly - pr + (uint32_t)i;
			for (uint64_t k = 0; k < ((uint64_t)1 << (pr - 1 - i)); k++) {
				const uint32_t t = J->tw[I][k];
				for (uint64_t l = 0; l < ((uint64_t)1 << i); l++) {
					uint64_t r0 = k << (i + 1) | l, r1 = r0 | (uint64_t)1 << i;
					uint32_t *u = J->data + (r0 << log_row) + 16 * J->begin, *v = J->data + (r1 << log_row) + 16 * J->begin;
					if (J->gfni) ntt_butterfly_run_gfni(u, v, J->end - J->begin, t, J->t2a, J->a2t);
					else ntt_butterfly_run_scalar(u, v, 16 * (J->end - J->begin), t);
				}
			}
		}
	} else {
		/* chunk c (a row of phase 1) = an independent transform of log_y' = ly - par_rounds layers with
		 * coset = c, coset_bits = par_rounds, i.e. block index c << (ly'-1-i) | k of the global layer i */
		const uint32_t pr = J->par_rounds, lyp = ly - pr, log_row = lx + lyp;
		const uint32_t skip2 = J->skip > pr ? J->skip - pr : 0;
		for (uint64_t c = J->begin; c < J->end; c++) {
			uint32_t *chunk = J->data + (c << log_row);
			for (int i = (int)lyp - 1 - (int)skip2; i >= 0; i--) {
				const uint64_t run = (uint64_t)1 << (lx + i);
				for (uint64_t k = 0; k < ((uint64_t)1 << (lyp - 1 - i)); k++) {
					const uint32_t t = J->tw[i][c << (lyp - 1 - i) | k];
					uint32_t *u = chunk + (k << (lx + i + 1)), *v = u + run;
					if (J->gfni) ntt_butterfly_run_gfni(u, v, run / 16, t, J->t2a, J->a2t);
					else ntt_butterfly_run_scalar(u, v, run, t);
				}
			}
		}
	}
	return NULL;
}

/* forward transform, shape (log_x, log_y, 0), coset 0; s = the oracle's s_evals (orc_ntt_s_evals, kt = 5) */
int cpu_ntt_forward(uint32_t *data, uint32_t log_x, uint32_t log_y, uint32_t skip, const u128u *s, uint32_t d, int n_threads, int use_gfni) {
	tower_init();
	if (log_x < 4 || log_y > d || skip > log_y) return 1;
	int gfni = use_gfni && cpu_has_gfni512();
	if (n_threads < 1) n_threads = 1;
	uint32_t log_thr = 0;
	while ((2u << log_thr) <= (uint32_t)n_threads) log_thr++;
	/* precomputed twiddles per layer (PrecomputedTwiddleAccess) */
	uint32_t **tw = malloc(sizeof(uint32_t *) * log_y);
	const uint32_t W = d - 1, row0 = d - log_y;
	for (uint32_t I = 0; I < log_y; I++) {
		uint64_t n = (uint64_t)1 << (log_y - 1 - I);
		tw[I] = malloc(sizeof(uint32_t) * n);
		tw[I][0] = 0;
		for (uint64_t j = 1; j < n; j++) {
			uint32_t b = (uint32_t)__builtin_ctzll(j);
			tw[I][j] = tw[I][j & (j - 1)] ^ (uint32_t)s[(row0 + I) * W + b];
		}
	}
	/* multithreaded.rs:139-152 with P = 16 x 32b: log_w = 4 */
	const uint32_t log_w = 4, total = log_x + log_y;
	uint32_t min_lw = log_w + 1 > log_x ? log_w + 1 : log_x;
	uint32_t log_height = total - min_lw < log_thr ? total - min_lw : log_thr;
	uint32_t log_width = total - (log_w + log_height); /* packed elements per row */
	uint32_t par_rounds = log_height;
	pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
	ntt_job *jobs = malloc(sizeof(ntt_job) * n_threads);
	for (int phase = 1; phase <= 2; phase++) {
		uint64_t units = phase == 1 ? (uint64_t)1 << log_width : (uint64_t)1 << par_rounds;
		if (phase == 1 && par_rounds <= skip) continue;
		int started = 0;
		for (int t = 0; t < n_threads; t++) {
			uint64_t b = units * t / n_threads, e = units * (t + 1) / n_threads;
			if (b == e) continue;
			jobs[started] = (ntt_job){data, log_x, log_y, skip, (const uint32_t *const *)tw, par_rounds, 0, b, e, phase, gfni,
									  affine_matrix(TOWER_TO_AES), affine_matrix(AES_TO_TOWER)};
			pthread_create(&th[started], NULL, ntt_worker, &jobs[started]);
			started++;
		}
		for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
	}
	for (uint32_t I = 0; I < log_y; I++) free(tw[I]);
	free(tw);
	free(th);
	free(jobs);
	return 0;
}

/* timed loop for bench.py: forward NTT of 2^(log_x+log_y) B32 coefficients, `reps` times; returns seconds */
double cpu_ntt_bench(uint32_t log_x, uint32_t log_y, uint32_t skip, int reps, int n_threads, const u128u *s, uint32_t d, int use_gfni) {
	uint64_t n = (uint64_t)1 << (log_x + log_y);
	uint32_t *buf = aligned_alloc(64, sizeof(uint32_t) * n);
	uint64_t x = 0x9876543;
	for (uint64_t i = 0; i < n; i++) {
		x = x * 6364136223846793005ull + 1442695040888963407ull;
		buf[i] = (uint32_t)(x >> 32);
	}
	cpu_ntt_forward(buf, log_x, log_y, skip, s, d, n_threads, use_gfni); /* warm-up, page-in */
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int r = 0; r < reps; r++) cpu_ntt_forward(buf, log_x, log_y, skip, s, d, n_threads, use_gfni);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	free(buf);
	return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* =================================================================================================
 * CPU arm of the bivariate-product sumcheck data plane (v3::BivariateSumcheckProver over FastCpuLayer,
 * core/src/protocols/sumcheck/v3/bivariate_product.rs:168-232, 303-408; the per-chunk kernel is
 * fast_compute/src/layer.rs:797-846): per round the two batched sums over all compositions, then the fold of
 * every multilinear.  Products on 4 x B128 per __m512i with the GFNI multiply; sums stay in the AES-tower basis
 * (the basis change is GF(2)-linear) and are converted once per thread.
 * ================================================================================================= */
typedef struct {
	u128u *const *mls;
	uint64_t begin, end, half;
	const uint32_t *ia, *ib;
	uint32_t n_comp;
	u128 *out; /* [2 * n_comp]: per composition sum at 1 and at infinity (tower basis) */
	int gfni;
} re_job;

TGT static void round_evals_gfni(re_job *J) {
	const __m512i t2a = _mm512_set1_epi64((long long)affine_matrix(TOWER_TO_AES));
	const __m512i a2t = _mm512_set1_epi64((long long)affine_matrix(AES_TO_TOWER));
	for (uint32_t c = 0; c < J->n_comp; c++) {
		const u128u *a = J->mls[J->ia[c]], *b = J->mls[J->ib[c]];
		__m512i s1 = _mm512_setzero_si512(), sinf = _mm512_setzero_si512();
		uint64_t i = J->begin;
		for (; i + 4 <= J->end; i += 4) {
			__m512i alo = _mm512_loadu_si512((const void *)(a + i)), ahi = _mm512_loadu_si512((const void *)(a + J->half + i));
			__m512i blo = _mm512_loadu_si512((const void *)(b + i)), bhi = _mm512_loadu_si512((const void *)(b + J->half + i));
			__m512i xa = _mm512_gf2p8affine_epi64_epi8(ahi, t2a, 0), xb = _mm512_gf2p8affine_epi64_epi8(bhi, t2a, 0);
			__m512i ya = _mm512_gf2p8affine_epi64_epi8(_mm512_xor_si512(alo, ahi), t2a, 0);
			__m512i yb = _mm512_gf2p8affine_epi64_epi8(_mm512_xor_si512(blo, bhi), t2a, 0);
			s1 = _mm512_xor_si512(s1, aes_mul7(xa, xb));
			sinf = _mm512_xor_si512(sinf, aes_mul7(ya, yb));
		}
		u128 lanes[4], r1 = 0, rinf = 0;
		_mm512_storeu_si512((void *)lanes, _mm512_gf2p8affine_epi64_epi8(s1, a2t, 0));
		r1 = lanes[0] ^ lanes[1] ^ lanes[2] ^ lanes[3];
		_mm512_storeu_si512((void *)lanes, _mm512_gf2p8affine_epi64_epi8(sinf, a2t, 0));
		rinf = lanes[0] ^ lanes[1] ^ lanes[2] ^ lanes[3];
		for (; i < J->end; i++) {
			r1 ^= b128_mul(a[J->half + i], b[J->half + i]);
			rinf ^= b128_mul(a[i] ^ a[J->half + i], b[i] ^ b[J->half + i]);
		}
		J->out[2 * c] = r1;
		J->out[2 * c + 1] = rinf;
	}
}
static void *re_worker(void *p) {
	re_job *J = (re_job *)p;
	if (J->gfni) {
		round_evals_gfni(J);
		return NULL;
	}
	for (uint32_t c = 0; c < J->n_comp; c++) {
		const u128u *a = J->mls[J->ia[c]], *b = J->mls[J->ib[c]];
		u128 r1 = 0, rinf = 0;
		for (uint64_t i = J->begin; i < J->end; i++) {
			r1 ^= b128_mul(a[J->half + i], b[J->half + i]);
			rinf ^= b128_mul(a[i] ^ a[J->half + i], b[i] ^ b[J->half + i]);
		}
		J->out[2 * c] = r1;
		J->out[2 * c + 1] = rinf;
	}
	return NULL;
}

/* [y_1, y_inf] of one round over m multilinears of 2^n_vars elements (bivariate_product.rs:303-408) */
int cpu_bivariate_round_evals(u128u *const *mls, uint32_t n_vars, const uint32_t *ia, const uint32_t *ib, uint32_t n_comp,
							  const u128u *batch_coeff, int n_threads, int use_gfni, u128u *out2) {
	tower_init();
	int gfni = use_gfni && cpu_has_gfni512();
	uint64_t half = (uint64_t)1 << (n_vars - 1);
	if (n_threads < 1 || half < (1u << 12)) n_threads = 1;
	pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
	re_job *jobs = malloc(sizeof(re_job) * n_threads);
	u128 *parts = calloc((size_t)2 * n_comp * n_threads, sizeof(u128));
	int started = 0;
	for (int t = 0; t < n_threads; t++) {
		uint64_t b = (half * t / n_threads) & ~3ull, e = t + 1 == n_threads ? half : (half * (t + 1) / n_threads) & ~3ull;
		if (b >= e) continue;
		jobs[started] = (re_job){mls, b, e, half, ia, ib, n_comp, parts + (size_t)2 * n_comp * started, gfni};
		if (n_threads == 1) re_worker(&jobs[started]);
		else pthread_create(&th[started], NULL, re_worker, &jobs[started]);
		started++;
	}
	if (n_threads > 1)
		for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
	u128 y1 = 0, yinf = 0, pw = 1;
	for (uint32_t c = 0; c < n_comp; c++) {
		u128 s1 = 0, sinf = 0;
		for (int t = 0; t < started; t++) {
			s1 ^= parts[(size_t)2 * n_comp * t + 2 * c];
			sinf ^= parts[(size_t)2 * n_comp * t + 2 * c + 1];
		}
		y1 ^= b128_mul(s1, pw);
		yinf ^= b128_mul(sinf, pw);
		pw = b128_mul(pw, *batch_coeff);
	}
	out2[0] = y1;
	out2[1] = yinf;
	free(th);
	free(jobs);
	free(parts);
	return gfni;
}

/* timed loop for bench.py: the whole sumcheck (n_vars rounds of round evaluations + fold of every multilinear)
 * over m multilinears and n_comp index pairs, `reps` times; returns seconds.  Writes a checksum of the round values. */
double cpu_bivariate_sumcheck_bench(uint32_t m, uint32_t n_vars, uint32_t n_comp, int reps, int n_threads, int use_gfni, u128u *checksum) {
	uint64_t n = (uint64_t)1 << n_vars;
	u128u **src = malloc(sizeof(u128u *) * m), **buf = malloc(sizeof(u128u *) * m);
	uint64_t s = 0x7777777;
	for (uint32_t t = 0; t < m; t++) {
		src[t] = aligned_alloc(64, sizeof(u128) * n);
		buf[t] = aligned_alloc(64, sizeof(u128) * n);
		for (uint64_t i = 0; i < n; i++) {
			s = s * 6364136223846793005ull + 1442695040888963407ull;
			src[t][i] = ((u128)s << 64) | (s * 0x9E3779B97F4A7C15ull);
		}
	}
	uint32_t *ia = malloc(4 * n_comp), *ib = malloc(4 * n_comp);
	for (uint32_t c = 0; c < n_comp; c++) {
		ia[c] = (c * 5 + 1) % m;
		ib[c] = (c * 3 + 2) % m;
	}
	u128 alpha = ((u128)0x0123456789ABCDEFull << 64) | 0x0F1E2D3C4B5A6978ull, z = ((u128)0x2E895399AF449ACEull << 64) | 0x499596F6E5FCCAFAull;
	u128 acc = 0;
	double total = 0;
	for (int r = -1; r < reps; r++) {
		for (uint32_t t = 0; t < m; t++) memcpy(buf[t], src[t], sizeof(u128) * n);
		struct timespec t0, t1;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		for (uint32_t v = n_vars; v >= 1; v--) {
			u128 y[2];
			cpu_bivariate_round_evals(buf, v, ia, ib, n_comp, (const u128u *)&alpha, n_threads, use_gfni, (u128u *)y);
			acc ^= y[0] ^ (y[1] << 1);
			uint64_t half = (uint64_t)1 << (v - 1);
			for (uint32_t t = 0; t < m; t++) cpu_fold(buf[t], buf[t] + half, half, (const u128u *)&z, half >= (1u << 13) ? n_threads : 1, use_gfni);
			z = z * 3 + 1;
		}
		clock_gettime(CLOCK_MONOTONIC, &t1);
		if (r >= 0) total += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
	}
	if (checksum) *checksum = acc;
	for (uint32_t t = 0; t < m; t++) {
		free(src[t]);
		free(buf[t]);
	}
	free(src);
	free(buf);
	free(ia);
	free(ib);
	return total;
}

/* =================================================================================================
 * CPU arm of the zerocheck multilinear rounds of the keccak chi constraints (keccak replay, phase
 * "zerocheck_rounds"): port of the eq-ind evaluator's hot loop (hal/src/sumcheck_round_calculation.rs:222-297
 * with core/src/protocols/sumcheck/prove/eq_ind.rs:646-731): per round and constraint
 *     C = out - (b0 + (b1 - 1) * b2)      at 1:        sum_i E[i] * C(hi[i])
 *                                         at infinity: sum_i E[i] * (b1' * b2')[i],  x' = hi - lo  (leading term)
 * then the fold of every multilinear and the halving of the eq-indicator.  GFNI multiply on 4 x B128 per register;
 * operands are converted to the AES-tower basis once per (column, chunk) the way the reference's batch_evaluate
 * works on packed slices.  Constraint c reads out = col[c], b_k = col[n_out + (c + k) % n_b].
 * ================================================================================================= */
typedef struct {
	u128u *const *cols;
	const u128u *eq;
	uint64_t begin, end, half;
	uint32_t n_out, n_b;
	int with_eval_1;
	u128 *out; /* [2 * n_out] */
} chi_job;

TGT static void *chi_worker_gfni(void *p) {
	chi_job *J = (chi_job *)p;
	const __m512i t2a = _mm512_set1_epi64((long long)affine_matrix(TOWER_TO_AES));
	const __m512i a2t = _mm512_set1_epi64((long long)affine_matrix(AES_TO_TOWER));
	const __m512i one = _mm512_gf2p8affine_epi64_epi8(_mm512_broadcast_i32x4(_mm_set_epi32(0, 0, 0, 1)), t2a, 0);
	for (uint32_t c = 0; c < J->n_out; c++) {
		const u128u *o = J->cols[c], *b0 = J->cols[J->n_out + c % J->n_b], *b1 = J->cols[J->n_out + (c + 1) % J->n_b], *b2 = J->cols[J->n_out + (c + 2) % J->n_b];
		__m512i s1 = _mm512_setzero_si512(), sinf = _mm512_setzero_si512();
		for (uint64_t i = J->begin; i + 4 <= J->end; i += 4) {
			const uint64_t h = J->half + i;
			__m512i e = _mm512_gf2p8affine_epi64_epi8(_mm512_loadu_si512((const void *)(J->eq + i)), t2a, 0);
			__m512i b1h = _mm512_loadu_si512((const void *)(b1 + h)), b2h = _mm512_loadu_si512((const void *)(b2 + h));
			__m512i b1p = _mm512_gf2p8affine_epi64_epi8(_mm512_xor_si512(b1h, _mm512_loadu_si512((const void *)(b1 + i))), t2a, 0);
			__m512i b2p = _mm512_gf2p8affine_epi64_epi8(_mm512_xor_si512(b2h, _mm512_loadu_si512((const void *)(b2 + i))), t2a, 0);
			sinf = _mm512_xor_si512(sinf, aes_mul7(e, aes_mul7(b1p, b2p)));
			if (J->with_eval_1) {
				__m512i x1 = _mm512_xor_si512(_mm512_gf2p8affine_epi64_epi8(b1h, t2a, 0), one), x2 = _mm512_gf2p8affine_epi64_epi8(b2h, t2a, 0);
				__m512i lin = _mm512_gf2p8affine_epi64_epi8(_mm512_xor_si512(_mm512_loadu_si512((const void *)(o + h)), _mm512_loadu_si512((const void *)(b0 + h))), t2a, 0);
				s1 = _mm512_xor_si512(s1, aes_mul7(e, _mm512_xor_si512(lin, aes_mul7(x1, x2))));
			}
		}
		u128 lanes[4];
		_mm512_storeu_si512((void *)lanes, _mm512_gf2p8affine_epi64_epi8(s1, a2t, 0));
		J->out[2 * c] = lanes[0] ^ lanes[1] ^ lanes[2] ^ lanes[3];
		_mm512_storeu_si512((void *)lanes, _mm512_gf2p8affine_epi64_epi8(sinf, a2t, 0));
		J->out[2 * c + 1] = lanes[0] ^ lanes[1] ^ lanes[2] ^ lanes[3];
	}
	return NULL;
}
static void *chi_worker_scalar(void *p) {
	chi_job *J = (chi_job *)p;
	for (uint32_t c = 0; c < J->n_out; c++) {
		const u128u *o = J->cols[c], *b0 = J->cols[J->n_out + c % J->n_b], *b1 = J->cols[J->n_out + (c + 1) % J->n_b], *b2 = J->cols[J->n_out + (c + 2) % J->n_b];
		u128 s1 = 0, sinf = 0;
		for (uint64_t i = J->begin; i < J->end; i++) {
			const uint64_t h = J->half + i;
			sinf ^= b128_mul(J->eq[i], b128_mul(b1[h] ^ b1[i], b2[h] ^ b2[i]));
			if (J->with_eval_1) s1 ^= b128_mul(J->eq[i], o[h] ^ b0[h] ^ b128_mul(b1[h] ^ 1, b2[h]));
		}
		J->out[2 * c] = s1;
		J->out[2 * c + 1] = sinf;
	}
	return NULL;
}

/* round values [constraint][at 1, at infinity] of one round; half a multiple of 4 or below 4 (scalar) */
int cpu_chi_round_evals(u128u *const *cols, uint32_t n_out, uint32_t n_b, uint32_t n_vars, const u128u *eq, int with_eval_1, int n_threads,
						int use_gfni, u128u *out /* 2 * n_out */) {
	tower_init();
	uint64_t half = (uint64_t)1 << (n_vars - 1);
	int gfni = use_gfni && cpu_has_gfni512() && half >= 4;
	if (n_threads < 1 || half < (1u << 10)) n_threads = 1;
	pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
	chi_job *jobs = malloc(sizeof(chi_job) * n_threads);
	u128 *parts = calloc((size_t)2 * n_out * n_threads, sizeof(u128));
	int started = 0;
	for (int t = 0; t < n_threads; t++) {
		uint64_t b = (half * t / n_threads) & ~3ull, e = t + 1 == n_threads ? half : (half * (t + 1) / n_threads) & ~3ull;
		if (b >= e) continue;
		jobs[started] = (chi_job){cols, eq, b, e, half, n_out, n_b, with_eval_1, parts + (size_t)2 * n_out * started};
		if (n_threads == 1) (gfni ? chi_worker_gfni : chi_worker_scalar)(&jobs[started]);
		else pthread_create(&th[started], NULL, gfni ? chi_worker_gfni : chi_worker_scalar, &jobs[started]);
		started++;
	}
	if (n_threads > 1)
		for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
	for (uint32_t k = 0; k < 2 * n_out; k++) {
		u128 s = 0;
		for (int t = 0; t < started; t++) s ^= parts[(size_t)2 * n_out * t + k];
		out[k] = s;
	}
	free(th);
	free(jobs);
	free(parts);
	return gfni;
}

/* timed loop: all n_vars rounds (round values, fold of the n_out + n_b multilinears, eq halving), `reps` times */
double cpu_chi_zerocheck_bench(uint32_t n_out, uint32_t n_b, uint32_t n_vars, int reps, int n_threads, int use_gfni, u128u *checksum) {
	const uint32_t m = n_out + n_b;
	uint64_t n = (uint64_t)1 << n_vars;
	u128u **buf = malloc(sizeof(u128u *) * m);
	u128u *eq = aligned_alloc(64, sizeof(u128) * (n / 2 > 4 ? n / 2 : 4));
	u128 *vals = malloc(sizeof(u128) * 2 * n_out);
	uint64_t s = 0x5151515;
	for (uint32_t t = 0; t < m; t++) buf[t] = aligned_alloc(64, sizeof(u128) * n);
	u128 z = ((u128)0x2E895399AF449ACEull << 64) | 0x499596F6E5FCCAFAull, acc = 0;
	double total = 0;
	for (int r = -1; r < reps; r++) {
		for (uint32_t t = 0; t < m; t++)
			for (uint64_t i = 0; i < n; i++) {
				s = s * 6364136223846793005ull + 1442695040888963407ull;
				buf[t][i] = ((u128)s << 64) | (s * 0x9E3779B97F4A7C15ull);
			}
		for (uint64_t i = 0; i < n / 2; i++) {
			s = s * 6364136223846793005ull + 1442695040888963407ull;
			eq[i] = ((u128)s << 64) | (s * 0x9E3779B97F4A7C15ull);
		}
		struct timespec t0, t1;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		for (uint32_t v = n_vars; v >= 1; v--) {
			cpu_chi_round_evals(buf, n_out, n_b, v, eq, v != n_vars, n_threads, use_gfni, (u128u *)vals);
			for (uint32_t k = 0; k < 2 * n_out; k++) acc ^= vals[k];
			uint64_t half = (uint64_t)1 << (v - 1);
			for (uint32_t t = 0; t < m; t++) cpu_fold(buf[t], buf[t] + half, half, (const u128u *)&z, half >= (1u << 13) ? n_threads : 1, use_gfni);
			for (uint64_t i = 0; i < half / 2; i++) eq[i] ^= eq[half / 2 + i]; /* fold_partial_eq_ind: E'[i] = E[i] + E[half + i] */
			z = z * 3 + 1;
		}
		clock_gettime(CLOCK_MONOTONIC, &t1);
		if (r >= 0) total += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
	}
	if (checksum) *checksum = acc;
	for (uint32_t t = 0; t < m; t++) free(buf[t]);
	free(buf);
	free(eq);
	free(vals);
	return total;
}

/* =================================================================================================
 * CPU arm of BASELINE config #3: zerocheck rounds of the u32_add gadget (m3/src/gadgets/add.rs; columns
 * x, y, cin, cout, z as 5 multilinears of n_vars variables; compositions C1 = (x + cin)(y + cin) + cin - cout and
 * C2 = x + y + cin - z).  Per round (eq-ind evaluator, hal/src/sumcheck_round_calculation.rs:222-297 with
 * core/src/protocols/sumcheck/prove/eq_ind.rs:646-731):
 *     at 1:        sum_i E[i] * C1(hi[i]),  sum_i E[i] * C2(hi[i])            (not in the first round)
 *     at infinity: sum_i E[i] * ((x' + cin')(y' + cin'))[i],  v' = hi - lo     (C2 is linear: no leading term)
 * then the fold of the 5 multilinears and the halving of the eq-indicator.  Same GFNI structure as the chi arm.
 * out = {C1 at 1, C1 at infinity, C2 at 1}
 * ================================================================================================= */
typedef struct {
	u128u *const *cols;
	const u128u *eq;
	uint64_t begin, end, half;
	int with_eval_1;
	u128 *out; /* [3] */
} add_job;

TGT static void *add_worker_gfni(void *p) {
	add_job *J = (add_job *)p;
	const __m512i t2a = _mm512_set1_epi64((long long)affine_matrix(TOWER_TO_AES));
	const __m512i a2t = _mm512_set1_epi64((long long)affine_matrix(AES_TO_TOWER));
	const u128u *x = J->cols[0], *y = J->cols[1], *ci = J->cols[2], *co = J->cols[3], *z = J->cols[4];
	__m512i s1 = _mm512_setzero_si512(), s2 = _mm512_setzero_si512(), sinf = _mm512_setzero_si512();
	for (uint64_t i = J->begin; i + 4 <= J->end; i += 4) {
		const uint64_t h = J->half + i;
		__m512i e = _mm512_gf2p8affine_epi64_epi8(_mm512_loadu_si512((const void *)(J->eq + i)), t2a, 0);
		__m512i xh = _mm512_loadu_si512((const void *)(x + h)), yh = _mm512_loadu_si512((const void *)(y + h));
		__m512i ch = _mm512_loadu_si512((const void *)(ci + h));
		__m512i cp = _mm512_xor_si512(ch, _mm512_loadu_si512((const void *)(ci + i)));
		__m512i xp = _mm512_gf2p8affine_epi64_epi8(_mm512_xor_si512(_mm512_xor_si512(xh, _mm512_loadu_si512((const void *)(x + i))), cp), t2a, 0);
		__m512i yp = _mm512_gf2p8affine_epi64_epi8(_mm512_xor_si512(_mm512_xor_si512(yh, _mm512_loadu_si512((const void *)(y + i))), cp), t2a, 0);
		sinf = _mm512_xor_si512(sinf, aes_mul7(e, aes_mul7(xp, yp)));
		if (J->with_eval_1) {
			__m512i a = _mm512_gf2p8affine_epi64_epi8(_mm512_xor_si512(xh, ch), t2a, 0), b = _mm512_gf2p8affine_epi64_epi8(_mm512_xor_si512(yh, ch), t2a, 0);
			__m512i lin = _mm512_gf2p8affine_epi64_epi8(_mm512_xor_si512(ch, _mm512_loadu_si512((const void *)(co + h))), t2a, 0);
			s1 = _mm512_xor_si512(s1, aes_mul7(e, _mm512_xor_si512(lin, aes_mul7(a, b))));
			__m512i l2 = _mm512_xor_si512(_mm512_xor_si512(xh, yh), _mm512_xor_si512(ch, _mm512_loadu_si512((const void *)(z + h))));
			s2 = _mm512_xor_si512(s2, aes_mul7(e, _mm512_gf2p8affine_epi64_epi8(l2, t2a, 0)));
		}
	}
	u128 lanes[4];
	_mm512_storeu_si512((void *)lanes, _mm512_gf2p8affine_epi64_epi8(s1, a2t, 0));
	J->out[0] = lanes[0] ^ lanes[1] ^ lanes[2] ^ lanes[3];
	_mm512_storeu_si512((void *)lanes, _mm512_gf2p8affine_epi64_epi8(sinf, a2t, 0));
	J->out[1] = lanes[0] ^ lanes[1] ^ lanes[2] ^ lanes[3];
	_mm512_storeu_si512((void *)lanes, _mm512_gf2p8affine_epi64_epi8(s2, a2t, 0));
	J->out[2] = lanes[0] ^ lanes[1] ^ lanes[2] ^ lanes[3];
	return NULL;
}
static void *add_worker_scalar(void *p) {
	add_job *J = (add_job *)p;
	const u128u *x = J->cols[0], *y = J->cols[1], *ci = J->cols[2], *co = J->cols[3], *z = J->cols[4];
	u128 s1 = 0, s2 = 0, sinf = 0;
	for (uint64_t i = J->begin; i < J->end; i++) {
		const uint64_t h = J->half + i;
		sinf ^= b128_mul(J->eq[i], b128_mul(x[h] ^ x[i] ^ ci[h] ^ ci[i], y[h] ^ y[i] ^ ci[h] ^ ci[i]));
		if (J->with_eval_1) {
			s1 ^= b128_mul(J->eq[i], b128_mul(x[h] ^ ci[h], y[h] ^ ci[h]) ^ ci[h] ^ co[h]);
			s2 ^= b128_mul(J->eq[i], x[h] ^ y[h] ^ ci[h] ^ z[h]);
		}
	}
	J->out[0] = s1, J->out[1] = sinf, J->out[2] = s2;
	return NULL;
}

int cpu_u32add_round_evals(u128u *const *cols, uint32_t n_vars, const u128u *eq, int with_eval_1, int n_threads, int use_gfni, u128u *out /* 3 */) {
	tower_init();
	uint64_t half = (uint64_t)1 << (n_vars - 1);
	int gfni = use_gfni && cpu_has_gfni512() && half >= 4;
	if (n_threads < 1 || half < (1u << 12)) n_threads = 1;
	pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
	add_job *jobs = malloc(sizeof(add_job) * n_threads);
	u128 *parts = calloc((size_t)3 * n_threads, sizeof(u128));
	int started = 0;
	for (int t = 0; t < n_threads; t++) {
		uint64_t b = (half * t / n_threads) & ~3ull, e = t + 1 == n_threads ? half : (half * (t + 1) / n_threads) & ~3ull;
		if (b >= e) continue;
		jobs[started] = (add_job){cols, eq, b, e, half, with_eval_1, parts + (size_t)3 * started};
		if (n_threads == 1) (gfni ? add_worker_gfni : add_worker_scalar)(&jobs[started]);
		else pthread_create(&th[started], NULL, gfni ? add_worker_gfni : add_worker_scalar, &jobs[started]);
		started++;
	}
	if (n_threads > 1)
		for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
	for (uint32_t k = 0; k < 3; k++) {
		u128 s = 0;
		for (int t = 0; t < started; t++) s ^= parts[(size_t)3 * t + k];
		out[k] = s;
	}
	free(th);
	free(jobs);
	free(parts);
	return gfni;
}

/* timed loop: all n_vars rounds (round values, fold of the 5 multilinears, eq halving), `reps` times */
double cpu_u32add_zerocheck_bench(uint32_t n_vars, int reps, int n_threads, int use_gfni, u128u *checksum) {
	const uint32_t m = 5;
	uint64_t n = (uint64_t)1 << n_vars;
	u128u *buf[5];
	u128u *eq = aligned_alloc(64, sizeof(u128) * (n / 2 > 4 ? n / 2 : 4));
	u128 vals[3];
	uint64_t s = 0x3232323;
	for (uint32_t t = 0; t < m; t++) buf[t] = aligned_alloc(64, sizeof(u128) * n);
	u128 z = ((u128)0x2E895399AF449ACEull << 64) | 0x499596F6E5FCCAFAull, acc = 0;
	double total = 0;
	for (int r = -1; r < reps; r++) {
		for (uint32_t t = 0; t < m; t++)
			for (uint64_t i = 0; i < n; i++) {
				s = s * 6364136223846793005ull + 1442695040888963407ull;
				buf[t][i] = ((u128)s << 64) | (s * 0x9E3779B97F4A7C15ull);
			}
		for (uint64_t i = 0; i < n / 2; i++) {
			s = s * 6364136223846793005ull + 1442695040888963407ull;
			eq[i] = ((u128)s << 64) | (s * 0x9E3779B97F4A7C15ull);
		}
		struct timespec t0, t1;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		for (uint32_t v = n_vars; v >= 1; v--) {
			cpu_u32add_round_evals(buf, v, eq, v != n_vars, n_threads, use_gfni, (u128u *)vals);
			acc ^= vals[0] ^ vals[1] ^ vals[2];
			uint64_t half = (uint64_t)1 << (v - 1);
			for (uint32_t t = 0; t < m; t++) cpu_fold(buf[t], buf[t] + half, half, (const u128u *)&z, half >= (1u << 13) ? n_threads : 1, use_gfni);
			for (uint64_t i = 0; i < half / 2; i++) eq[i] ^= eq[half / 2 + i];
			z = z * 3 + 1;
		}
		clock_gettime(CLOCK_MONOTONIC, &t1);
		if (r >= 0) total += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
	}
	if (checksum) *checksum = acc;
	for (uint32_t t = 0; t < m; t++) free(buf[t]);
	free(eq);
	return total;
}

/* =================================================================================================
 * CPU arm of the eq-indicator tensor expansion (crates/math/src/tensor_prod_eq_ind.rs:35-77, the FastCpuLayer's
 * tensor_expand, fast_compute/src/layer.rs): round r maps the 2^r filled elements to 2^(r+1):
 *     hi[i] = lo[i] * z_r;   lo[i] -= hi[i]
 * with the GFNI multiply on 4 x B128 per register, the rounds split over threads once they are large enough
 * (the reference parallelises the large rounds with rayon the same way).
 * ================================================================================================= */
typedef struct {
	u128u *lo, *hi;
	uint64_t n;
	u128 z;
	int gfni;
} exp_job;

TGT static void expand_gfni(u128u *lo, u128u *hi, uint64_t n, u128 z) {
	const __m512i t2a = _mm512_set1_epi64((long long)affine_matrix(TOWER_TO_AES));
	const __m512i a2t = _mm512_set1_epi64((long long)affine_matrix(AES_TO_TOWER));
	__m512i zv = _mm512_gf2p8affine_epi64_epi8(_mm512_broadcast_i32x4(_mm_loadu_si128((const __m128i *)&z)), t2a, 0);
	uint64_t i = 0;
	for (; i + 4 <= n; i += 4) {
		__m512i a = _mm512_loadu_si512((const void *)(lo + i));
		__m512i p = _mm512_gf2p8affine_epi64_epi8(aes_mul7(_mm512_gf2p8affine_epi64_epi8(a, t2a, 0), zv), a2t, 0);
		_mm512_storeu_si512((void *)(hi + i), p);
		_mm512_storeu_si512((void *)(lo + i), _mm512_xor_si512(a, p));
	}
	for (; i < n; i++) {
		u128 p = b128_mul(lo[i], z);
		hi[i] = p, lo[i] ^= p;
	}
}
static void *exp_worker(void *p) {
	exp_job *j = (exp_job *)p;
	if (j->gfni) expand_gfni(j->lo, j->hi, j->n, j->z);
	else
		for (uint64_t i = 0; i < j->n; i++) {
			u128 q = b128_mul(j->lo[i], j->z);
			j->hi[i] = q, j->lo[i] ^= q;
		}
	return NULL;
}
/* data[0 .. 2^log_n) filled; expands by the k coordinates to 2^(log_n + k) elements in place */
int cpu_tensor_expand(u128u *data, uint32_t log_n, const u128u *coords, uint32_t k, int n_threads, int use_gfni) {
	tower_init();
	int gfni = use_gfni && cpu_has_gfni512();
	if (n_threads < 1) n_threads = 1;
	pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
	exp_job *jobs = malloc(sizeof(exp_job) * n_threads);
	for (uint32_t r = 0; r < k; r++) {
		uint64_t n = (uint64_t)1 << (log_n + r);
		int nt = n >= (1u << 14) ? n_threads : 1;
		int started = 0;
		for (int t = 0; t < nt; t++) {
			uint64_t b = (n * t / nt) & ~3ull, e = t + 1 == nt ? n : (n * (t + 1) / nt) & ~3ull;
			if (b >= e) continue;
			jobs[started] = (exp_job){data + b, data + n + b, e - b, coords[r], gfni};
			if (nt == 1) exp_worker(&jobs[started]);
			else pthread_create(&th[started], NULL, exp_worker, &jobs[started]);
			started++;
		}
		if (nt > 1)
			for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
	}
	free(th);
	free(jobs);
	return gfni;
}
double cpu_tensor_expand_bench(uint32_t k, int reps, int n_threads, int use_gfni, u128u *checksum) {
	uint64_t n = (uint64_t)1 << k;
	u128u *data = aligned_alloc(64, sizeof(u128) * n);
	u128 *coords = malloc(sizeof(u128) * (k ? k : 1));
	uint64_t s = 0x7E7E7E7;
	double total = 0;
	u128 acc = 0;
	for (int r = -1; r < reps; r++) {
		for (uint32_t t = 0; t < k; t++) {
			s = s * 6364136223846793005ull + 1442695040888963407ull;
			coords[t] = ((u128)s << 64) | (s * 0x9E3779B97F4A7C15ull);
		}
		data[0] = 1;
		struct timespec t0, t1;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		cpu_tensor_expand(data, 0, (const u128u *)coords, k, n_threads, use_gfni);
		clock_gettime(CLOCK_MONOTONIC, &t1);
		if (r >= 0) total += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
		acc ^= data[n - 1] ^ data[n / 2];
	}
	if (checksum) *checksum = acc;
	free(data);
	free(coords);
	return total;
}
