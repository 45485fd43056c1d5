/*
 * oracle/univariate.c -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Specification of the zerocheck univariate-skip round, restated from the
 * reference's own naive check of `zerocheck_univariate_evals`
 * (core/src/protocols/sumcheck/prove/univariate.rs:860-915; the optimised routine is :235-500):
 *
 *   R[c][i] = sum_{s < 2^(n-k)} eq[s] * C_c( P_0(s, x_i), ..., P_{m-1}(s, x_i) ),   i < max_domain_size - 2^k
 *   x_i    = FDomain::from(2^k + i)                        (B8 element with that integer value)
 *   P_j(s, x) = extrapolation of the 2^k values M_j[s*2^k + t], t < 2^k, taken at the points
 *               BinarySubspace::with_dim(k).iter() = the B8 elements 0, 1, ..., 2^k - 1, to x
 *               (EvaluationDomain::extrapolate, math/src/univariate.rs:177-202: Lagrange form)
 * The reference evaluates each composition only on deg(C)*2^k points and extends to max_domain_size by
 * NTT interpolation (extrapolate_round_evals, univariate.rs:570-640); the round polynomial has degree
 * <= deg(C)*(2^k - 1), so evaluating the definition at every point gives the same values.
 * Sub-field scalars embed into B128 as the same integers (binary_field.rs:505-527), so everything is
 * computed in B128 arithmetic.  No product code implements this row yet: this file is its oracle.
 */
#include <stdint.h>
#include <stdlib.h>

#include "tower.h"

typedef u128 __attribute__((aligned(8))) u128u;
typedef struct {
	uint32_t op, l;
	uint64_t r;
	uint64_t c_lo, c_hi;
} orc_expr_step;

static u128 upow(u128 x, uint64_t e) {
	u128 r = 1;
	while (e) {
		if (e & 1) r = b128_mul(r, x);
		x = b128_mul(x, x);
		e >>= 1;
	}
	return r;
}
static u128 ueval(const orc_expr_step *steps, uint32_t n, const u128 *q, u128 *tmp) {
	if (!n) return 0;
	for (uint32_t s = 0; s < n; s++) {
		const orc_expr_step *st = &steps[s];
		switch (st->op) {
		case 0: tmp[s] = tmp[st->l] ^ tmp[st->r]; break;
		case 1: tmp[s] = b128_mul(tmp[st->l], tmp[st->r]); break;
		case 2: tmp[s] = upow(tmp[st->l], st->r); break;
		case 3: tmp[s] = ((u128)st->c_hi << 64) | st->c_lo; break;
		default: tmp[s] = q[st->l]; break;
		}
	}
	return tmp[n - 1];
}

/* Lagrange basis over the points 0..n-1 at x:  L_t(x) = prod_{u != t} (x - u) / (t - u)   (out: n values) */
static void lagrange_n(uint32_t n, u128 x, u128u *out) {
	for (uint32_t t = 0; t < n; t++) {
		u128 num = 1, den = 1;
		for (uint32_t u = 0; u < n; u++) {
			if (u == t) continue;
			num = b128_mul(num, x ^ (u128)u);
			den = b128_mul(den, (u128)t ^ (u128)u);
		}
		out[t] = b128_mul(num, tower_invert(den, 7));
	}
}

/*
 * extrapolate_round_evals (univariate.rs:565-640): the reference evaluates composition c only at the
 * (deg_c - 1) * 2^k points following the skipped domain, re-adds 2^k ZERO evaluations in front (an honest
 * prover's values there), interpolates those deg_c * 2^k values (OddInterpolate::inverse_transform,
 * ntt/src/odd_interpolate.rs:60-72: "the unique univariate polynomial P of degree less than d * 2^l" whose
 * evaluations on the first d * 2^l field elements are the data; then a forward NTT) and evaluates it on the rest
 * of the domain.  Restated as
 * Lagrange extrapolation.  vals: n_points = max_domain_size - 2^k entries of which the first
 * (deg - 1) * 2^k are inputs; the rest are overwritten.
 */
void orc_extrapolate_round_evals(uint32_t skip, uint32_t degree, uint32_t max_domain_size, u128u *vals) {
	tower_init();
	const uint32_t K = 1u << skip, n_points = max_domain_size - K;
	const uint32_t n = (degree ? degree : 1) * K; /* degree 0: no evaluations, 2^k zeros -> zero polynomial */
	if (n >= max_domain_size) return;
	u128u *lag = malloc(sizeof(u128) * n);
	for (uint32_t i = n - K; i < n_points; i++) {
		lagrange_n(n, (u128)(K + i), lag);
		u128 acc = 0;
		for (uint32_t t = K; t < n; t++) acc ^= b128_mul(lag[t], vals[t - K]);
		vals[i] = acc;
	}
	free(lag);
}

/* Lagrange basis over the points 0..2^k-1 at x  (out: 2^k values) */
void orc_lagrange_evals(uint32_t k, const u128u *x, u128u *out) {
	tower_init();
	uint32_t n = 1u << k;
	lagrange_n(n, *x, out);
}

/* scalar `idx` of a packed sub-field multilinear (tower level lvl, 2^(7-lvl) scalars per B128 word) */
static inline u128 sub_scalar(const u128u *packed, uint32_t lvl, uint64_t idx) {
	uint32_t per_log = 7 - lvl;
	u128 w = packed[idx >> per_log];
	uint32_t j = (uint32_t)(idx & ((1u << per_log) - 1));
	if (lvl == 7) return w;
	u128 mask = (((u128)1) << (1u << lvl)) - 1;
	return (w >> (j << lvl)) & mask;
}

/*
 * out[c * n_points + i], n_points = max_domain_size - 2^skip.  mls[j]: packed sub-field multilinear of
 * n_vars variables at tower level levels[j]; eq_ind: 2^(n_vars - skip) B128 values.
 */
int orc_zerocheck_univariate_evals(const u128u *const *mls, const uint32_t *levels, uint32_t m, uint32_t n_vars, uint32_t skip,
								   const u128u *eq_ind, const orc_expr_step *const *comps, const uint32_t *comp_steps, uint32_t n_comp,
								   uint32_t max_domain_size, u128u *out) {
	tower_init();
	if (skip > n_vars || skip > 8 || max_domain_size > 256 || max_domain_size < (1u << skip)) return 1;
	const uint32_t K = 1u << skip, n_points = max_domain_size - K;
	const uint64_t n_sub = (uint64_t)1 << (n_vars - skip);
	u128u *lag = malloc(sizeof(u128) * K);
	u128 *q = malloc(sizeof(u128) * (m + 1)), *tmp = malloc(sizeof(u128) * 256);
	for (uint32_t i = 0; i < n_points; i++) {
		u128u x = (u128)(K + i);
		orc_lagrange_evals(skip, &x, lag);
		for (uint32_t c = 0; c < n_comp; c++) out[c * n_points + i] = 0;
		for (uint64_t s = 0; s < n_sub; s++) {
			for (uint32_t j = 0; j < m; j++) {
				u128 p = 0;
				for (uint32_t t = 0; t < K; t++) p ^= b128_mul(lag[t], sub_scalar(mls[j], levels[j], s * K + t));
				q[j] = p;
			}
			for (uint32_t c = 0; c < n_comp; c++) out[c * n_points + i] ^= b128_mul(eq_ind[s], ueval(comps[c], comp_steps[c], q, tmp));
		}
	}
	free(lag);
	free(q);
	free(tmp);
	return 0;
}
