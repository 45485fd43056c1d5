/*
 * oracle/hal.c -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Scalar restatements of the old-HAL (binius_hal::ComputationBackend) hot-path functions, packed
 * width 1.  Paths relative to /root/reference.
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

static u128 pow128(u128 x, uint64_t e) {
	u128 r = 1;
	while (e) {
		if (e & 1) r = b128_mul(r, x);
		x = b128_mul(x, x);
		e >>= 1;
	}
	return r;
}
static u128 eval_expr(const orc_expr_step *steps, uint32_t n, const u128 *q, u128 *tmp) {
	if (!n) return 0;
	for (uint32_t s = 0; s < n; s++) {
		const orc_expr_step *st = &steps[s];
		switch (st->op) {
		case 0: tmp[s] = tmp[st->l] ^ tmp[st->r]; break;
		case 1: tmp[s] = b128_mul(tmp[st->l], tmp[st->r]); break;
		case 2: tmp[s] = pow128(tmp[st->l], st->r); break;
		case 3: tmp[s] = ((u128)st->c_hi << 64) | st->c_lo; break;
		default: tmp[s] = q[st->l]; break;
		}
	}
	return tmp[n - 1];
}

/*
 * math/src/fold.rs:648-696  fold_left_lerp_inplace (P::WIDTH = 1), as driven by
 * hal/src/sumcheck_folding.rs:223-237 (Folded branch of fold_multilinears_high_to_low):
 *   evals has `prefix` stored elements of a 2^log_n multilinear, the rest equal `suffix`;
 *   i <  pivot          : e[i] += (e[half+i] - e[i]) * z
 *   pivot <= i < upper  : e[i] += (suffix    - e[i]) * z
 *   truncate to `upper`.  Returns the new length.
 */
uint64_t orc_fold_left_lerp_inplace(u128u *evals, uint64_t prefix, const u128u *suffix, uint32_t log_n, const u128u *z) {
	tower_init();
	uint64_t half = (uint64_t)1 << (log_n - 1);
	uint64_t pivot = prefix > half ? prefix - half : 0;
	uint64_t upper = prefix < half ? prefix : half;
	for (uint64_t i = 0; i < pivot; i++) evals[i] ^= b128_mul(evals[half + i] ^ evals[i], *z);
	for (uint64_t i = pivot; i < upper; i++) evals[i] ^= b128_mul(*suffix ^ evals[i], *z);
	return upper;
}

/*
 * hal/src/sumcheck_round_calculation.rs:126-349 (calculate_round_evals_with_access, HighToLowAccess
 * :507-604) with the eq-ind evaluator of core/src/protocols/sumcheck/prove/eq_ind.rs:646-731:
 *   out[c*n_points + p] = sum_{i < 2^(n-1)} E[i] * C_c^{(p)}(P_0(i), ..., P_{m-1}(i))
 *   code 1: P = hi            ; code 2: P = hi - lo, evaluated on leading_term(C)  (:241-249)
 *   code >= 3: P = lo + (hi - lo) * point[p]                                       (:250-271)
 * multilinear t has lens[t] stored elements followed by the implicit constant suffix[t]
 * (Folded multilinears, :573-600).  The reference's const-suffix shortcut (eq_ind.rs:704-721) is an
 * analytic evaluation of the same sum, so the brute-force definition is the specification.
 */
int orc_eq_ind_round_evals(const u128u *const *mls, const uint64_t *lens, const u128u *suffix, uint32_t m,
						   uint32_t n_vars, const u128u *eq_ind, const orc_expr_step *const *comps,
						   const uint32_t *comp_steps, const orc_expr_step *const *leads,
						   const uint32_t *lead_steps, uint32_t n_comp, const uint32_t *codes,
						   const u128u *points, uint32_t n_points, u128u *out) {
	tower_init();
	uint64_t half = (uint64_t)1 << (n_vars - 1);
	u128 *q = malloc(sizeof(u128) * (m + 1)), *tmp = malloc(sizeof(u128) * 256);
	for (uint32_t c = 0; c < n_comp; c++)
		for (uint32_t p = 0; p < n_points; p++) {
			u128 acc = 0;
			for (uint64_t i = 0; i < half; i++) {
				for (uint32_t t = 0; t < m; t++) {
					u128 lo = i < lens[t] ? mls[t][i] : suffix[t];
					u128 hi = half + i < lens[t] ? mls[t][half + i] : suffix[t];
					if (codes[p] == 1) q[t] = hi;
					else if (codes[p] == 2) q[t] = hi ^ lo;
					else q[t] = lo ^ b128_mul(hi ^ lo, points[p]);
				}
				u128 v = codes[p] == 2 ? eval_expr(leads[c], lead_steps[c], q, tmp) : eval_expr(comps[c], comp_steps[c], q, tmp);
				acc ^= b128_mul(v, eq_ind[i]);
			}
			out[c * n_points + p] = acc;
		}
	free(q);
	free(tmp);
	return 0;
}

/* core/src/protocols/sumcheck/prove/common.rs:60-68  fold_partial_eq_ind (high-to-low): E'[i] = E[i] + E[half+i] */
void orc_fold_partial_eq_ind(u128u *e, uint64_t n) {
	for (uint64_t i = 0; i < n / 2; i++) e[i] ^= e[n / 2 + i];
}

/*
 * Both evaluation orders and both evaluator kinds of hal/src/sumcheck_round_calculation.rs:126-349:
 *   order 1 = HighToLowAccess (:507-604): (lo, hi) = (M[i], M[half + i])
 *   order 0 = LowToHighAccess (:408-504): (lo, hi) = (M[2i], M[2i + 1])   (interleaved, then unzipped)
 *   eq_ind != NULL: eq-ind evaluator (core/.../prove/eq_ind.rs:646-731), sums weighted by eq_ind[i]
 *   eq_ind == NULL: regular evaluator (core/.../prove/regular_sumcheck.rs:233-277), plain sums
 */
int orc_sumcheck_round_evals(uint32_t order, const u128u *const *mls, const uint64_t *lens, const u128u *suffix, uint32_t m,
							 uint32_t n_vars, const u128u *eq_ind, const orc_expr_step *const *comps,
							 const uint32_t *comp_steps, const orc_expr_step *const *leads,
							 const uint32_t *lead_steps, uint32_t n_comp, const uint32_t *codes,
							 const u128u *points, uint32_t n_points, u128u *out) {
	tower_init();
	uint64_t half = (uint64_t)1 << (n_vars - 1);
	u128 *q = malloc(sizeof(u128) * (m + 1)), *tmp = malloc(sizeof(u128) * 256);
	for (uint32_t c = 0; c < n_comp; c++)
		for (uint32_t p = 0; p < n_points; p++) {
			u128 acc = 0;
			for (uint64_t i = 0; i < half; i++) {
				uint64_t i_lo = order ? i : 2 * i, i_hi = order ? half + i : 2 * i + 1;
				for (uint32_t t = 0; t < m; t++) {
					u128 lo = i_lo < lens[t] ? mls[t][i_lo] : suffix[t];
					u128 hi = i_hi < lens[t] ? mls[t][i_hi] : suffix[t];
					if (codes[p] == 1) q[t] = hi;
					else if (codes[p] == 2) q[t] = hi ^ lo;
					else q[t] = lo ^ b128_mul(hi ^ lo, points[p]);
				}
				u128 v = codes[p] == 2 ? eval_expr(leads[c], lead_steps[c], q, tmp) : eval_expr(comps[c], comp_steps[c], q, tmp);
				acc ^= eq_ind ? b128_mul(v, eq_ind[i]) : v;
			}
			out[c * n_points + p] = acc;
		}
	free(q);
	free(tmp);
	return 0;
}

/*
 * math/src/fold.rs:528-575 fold_right_lerp (P::WIDTH = PE::WIDTH = 1) as driven by the Folded branch of
 * fold_multilinears_low_to_high (hal/src/sumcheck_folding.rs:117-143):
 *   out[i] = in[2i] + (in[2i+1] - in[2i]) * z        i < evals_size / 2
 *   odd evals_size: out[evals_size/2] = in[last] + (suffix - in[last]) * z
 * Returns the number of output elements, ceil(evals_size / 2).
 */
uint64_t orc_fold_right_lerp(const u128u *evals, uint64_t evals_size, const u128u *suffix, const u128u *z, u128u *out) {
	tower_init();
	uint64_t folded = evals_size >> 1;
	for (uint64_t i = 0; i < folded; i++) out[i] = evals[2 * i] ^ b128_mul(evals[2 * i + 1] ^ evals[2 * i], *z);
	if (evals_size & 1) out[folded] = evals[2 * folded] ^ b128_mul(*suffix ^ evals[2 * folded], *z);
	return (evals_size + 1) >> 1;
}

/* core/src/protocols/sumcheck/prove/common.rs:37-57 fold_partial_eq_ind (low-to-high): E'[i] = E[2i] + E[2i+1] */
void orc_fold_partial_eq_ind_low_to_high(const u128u *e, uint64_t n, u128u *out) {
	for (uint64_t i = 0; i < n / 2; i++) out[i] = e[2 * i] ^ e[2 * i + 1];
}
