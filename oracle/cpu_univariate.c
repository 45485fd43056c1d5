/*
 * oracle/cpu_univariate.c -- TEST / MEASUREMENT INFRASTRUCTURE ONLY (never on the product path).
 *
 * Multi-threaded, table-driven CPU arm of the zerocheck univariate-skip round for the B1 / degree <= 2 shape
 * (the timed CPU baseline next to k_uni_b8; kind "port": a restatement, not the Rust binary).  Same result
 * as oracle/univariate.c (checked in tests/test_oracle_univariate.py); the reference's own optimised routine
 * is core/src/protocols/sumcheck/prove/univariate.rs:235-500 (rayon over sub-cube batches, NTT extrapolation,
 * PackedSubfield composition evaluation).  Per sub-cube: Lagrange-form extrapolation of every column to all
 * points through byte-indexed tables (8 bits of the sub-cube per lookup, 8 points per 64-bit XOR), monomial
 * evaluation with the 64 KiB B8 product table, eq[s] * value through two 16-entry B128 tables.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "tower.h"

typedef u128 __attribute__((aligned(8))) u128u;
void orc_lagrange_evals(uint32_t k, const u128u *x, u128u *out);

typedef struct {
	const uint8_t *const *cols;
	uint32_t m, skip, n_pts, n_comp;
	const u128u *eq;
	const uint16_t *mono_a, *mono_b;
	const uint8_t *mono_c;
	const uint32_t *comp_first, *comp_cnt;
	const uint8_t *lut; /* [K/8][256][n_pts] */
	uint64_t s0, s1;
	u128 *acc; /* [n_comp][n_pts] */
} uni_job;

static void *uni_worker(void *arg) {
	uni_job *J = arg;
	const uint32_t K = 1u << J->skip, KB = K / 8, P = J->n_pts, PW = (P + 7) / 8;
	uint64_t *q = malloc((size_t)J->m * PW * 8);
	uint8_t *val = malloc(PW * 8);
	for (uint64_t s = J->s0; s < J->s1; s++) {
		for (uint32_t j = 0; j < J->m; j++) {
			uint64_t *qj = q + (size_t)j * PW;
			const uint8_t *bits = J->cols[j] + s * KB;
			memset(qj, 0, PW * 8);
			for (uint32_t b = 0; b < KB; b++) {
				const uint8_t *row = J->lut + ((size_t)b * 256 + bits[b]) * PW * 8;
				for (uint32_t w = 0; w < PW; w++) {
					uint64_t r;
					memcpy(&r, row + 8 * w, 8);
					qj[w] ^= r;
				}
			}
		}
		u128 elo[16], ehi[16];
		const u128 e = J->eq[s];
		for (uint32_t n = 0; n < 16; n++) {
			elo[n] = tower_mul(e, (u128)n, 7);
			ehi[n] = tower_mul(e, (u128)(n << 4), 7);
		}
		for (uint32_t c = 0; c < J->n_comp; c++) {
			memset(val, 0, PW * 8);
			for (uint32_t t = J->comp_first[c]; t < J->comp_first[c] + J->comp_cnt[c]; t++) {
				const uint16_t a = J->mono_a[t], b = J->mono_b[t];
				const uint8_t cf = J->mono_c[t];
				const uint8_t *qa = a == 0xFFFF ? NULL : (const uint8_t *)(q + (size_t)a * PW);
				const uint8_t *qb = b == 0xFFFF ? NULL : (const uint8_t *)(q + (size_t)b * PW);
				for (uint32_t i = 0; i < P; i++) {
					uint8_t v = cf;
					if (qa) {
						v = qa[i];
						if (qb) v = TOWER_MUL8[v][qb[i]];
						if (cf != 1) v = TOWER_MUL8[v][cf];
					}
					val[i] ^= v;
				}
			}
			u128 *ac = J->acc + (size_t)c * P;
			for (uint32_t i = 0; i < P; i++) ac[i] ^= elo[val[i] & 15] ^ ehi[val[i] >> 4];
		}
	}
	free(q);
	free(val);
	return NULL;
}

/*
 * cols[j]: the packed B1 column (2^n_vars bits, little-endian bytes); every composition is a list of monomials
 * coef * x_a * x_b (a, b = column index or 0xFFFF) and is evaluated at the n_pts points following the skipped
 * domain.  out[c * n_pts + i].  skip >= 3.
 */
int orc_cpu_univariate_b1(const uint8_t *const *cols, uint32_t m, uint32_t n_vars, uint32_t skip, const u128u *eq, const uint16_t *mono_a,
						  const uint16_t *mono_b, const uint8_t *mono_c, const uint32_t *comp_first, const uint32_t *comp_cnt, uint32_t n_comp,
						  uint32_t n_pts, uint32_t n_threads, u128u *out) {
	tower_init();
	if (skip < 3 || skip > 7 || skip > n_vars || n_pts + (1u << skip) > 256) return 1;
	const uint32_t K = 1u << skip, KB = K / 8, PW = (n_pts + 7) / 8;
	const uint64_t n_sub = (uint64_t)1 << (n_vars - skip);
	if (n_threads < 1) n_threads = 1;
	if (n_threads > n_sub) n_threads = (uint32_t)n_sub;
	/* lag[i][t] in B8, then the byte-indexed tables */
	/* L_t(x) = prod_{u != t} (x - u) / (t - u) with the 8-bit tables (all points are B8 elements);
	 * tests/test_oracle_univariate.py checks the result against orc_lagrange_evals through the whole round */
	uint8_t *lag = malloc((size_t)n_pts * K);
	u128u *tmp = malloc(sizeof(u128) * K);
	for (uint32_t i = 0; i < n_pts; i++) {
		const uint8_t x = (uint8_t)(K + i);
		for (uint32_t t = 0; t < K; t++) {
			uint8_t num = 1, den = 1;
			for (uint32_t u = 0; u < K; u++) {
				if (u == t) continue;
				num = TOWER_MUL8[num][x ^ u];
				den = TOWER_MUL8[den][t ^ u];
			}
			lag[(size_t)i * K + t] = TOWER_MUL8[num][TOWER_INV8[den]];
		}
	}
	uint8_t *lut = calloc((size_t)KB * 256 * PW * 8, 1);
	for (uint32_t b = 0; b < KB; b++)
		for (uint32_t pat = 0; pat < 256; pat++) {
			uint8_t *row = lut + ((size_t)b * 256 + pat) * PW * 8;
			for (uint32_t i = 0; i < n_pts; i++) {
				uint8_t v = 0;
				for (uint32_t bit = 0; bit < 8; bit++)
					if (pat >> bit & 1) v ^= lag[(size_t)i * K + 8 * b + bit];
				row[i] = v;
			}
		}
	pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
	uni_job *jobs = malloc(sizeof(uni_job) * n_threads);
	const uint64_t per = (n_sub + n_threads - 1) / n_threads;
	for (uint32_t t = 0; t < n_threads; t++) {
		uni_job *J = &jobs[t];
		J->cols = cols, J->m = m, J->skip = skip, J->n_pts = n_pts, J->n_comp = n_comp, J->eq = eq;
		J->mono_a = mono_a, J->mono_b = mono_b, J->mono_c = mono_c, J->comp_first = comp_first, J->comp_cnt = comp_cnt, J->lut = lut;
		J->s0 = t * per < n_sub ? t * per : n_sub;
		J->s1 = (t + 1) * per < n_sub ? (t + 1) * per : n_sub;
		J->acc = calloc((size_t)n_comp * n_pts, sizeof(u128));
		pthread_create(&th[t], NULL, uni_worker, J);
	}
	for (size_t k = 0; k < (size_t)n_comp * n_pts; k++) out[k] = 0;
	for (uint32_t t = 0; t < n_threads; t++) {
		pthread_join(th[t], NULL);
		for (size_t k = 0; k < (size_t)n_comp * n_pts; k++) out[k] ^= jobs[t].acc[k];
		free(jobs[t].acc);
	}
	free(th), free(jobs), free(lut), free(lag), free(tmp);
	return 0;
}
