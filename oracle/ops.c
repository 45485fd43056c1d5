/*
 * oracle/ops.c -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Scalar, single-thread CPU restatement of the reference's ComputeLayer ops, fold, round-evaluation
 * and additive-NTT algorithms.  Each function cites the reference file:line it follows (paths
 * relative to /root/reference).  The reference is Rust and cannot be built in this image (no
 * rustc/cargo), so this file is a restatement pinned by the field KATs (tests/test_oracle_field.py)
 * plus the algebraic properties the reference's own tests assert (tests/test_oracle_ops.py).
 * The reference stores no input->output vectors for these ops, so op-level parity
 * with the Rust binary is pinned transitively: exact field arithmetic + restated loops.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "tower.h"

/* u128 arrays coming from numpy are only guaranteed 8-byte aligned */
typedef u128 __attribute__((aligned(8))) u128u;

#define ORC_OK 0
#define ORC_INPUT_VALIDATION 1

/* ---------------------------------------------------------------------------------------------
 * compute/src/cpu/layer.rs:393-408   extrapolate_line:  e0[i] += (e1[i] - e0[i]) * z
 * ------------------------------------------------------------------------------------------- */
int orc_extrapolate_line(u128u *e0, const u128u *e1, uint64_t n, const u128u *z) {
	tower_init();
	for (uint64_t i = 0; i < n; i++) e0[i] ^= b128_mul(e1[i] ^ e0[i], *z);
	return ORC_OK;
}

/* compute/src/cpu/layer.rs:282-302   tensor_expand (note: += into the upper half) */
int orc_tensor_expand(u128u *data, uint64_t data_len, uint32_t log_n, const u128u *coords, uint32_t k) {
	tower_init();
	if (data_len != ((uint64_t)1 << (log_n + k))) return ORC_INPUT_VALIDATION;
	for (uint32_t i = 0; i < k; i++) {
		uint64_t half = (uint64_t)1 << (log_n + i);
		for (uint64_t j = 0; j < half; j++) {
			u128 p = b128_mul(data[j], coords[i]);
			data[j] ^= p;
			data[half + j] ^= p;
		}
	}
	return ORC_OK;
}

/* limb j (2^lvl bits) of a B128 -- binary_field.rs:600-607, 628-657 (iter_bases: low limb first) */
static inline u128 limb(u128 a, uint32_t lvl, uint32_t j) {
	if (lvl == 7) return a;
	uint32_t w = 1u << lvl;
	return (a >> (j * w)) & ((((u128)1) << w) - 1);
}

/* compute/src/cpu/layer.rs:205-236   inner_product(SubfieldSlice{a, lvl}, b) */
int orc_inner_product(const u128u *a, uint64_t n_a, uint32_t lvl, const u128u *b, uint64_t n_b, u128u *out) {
	tower_init();
	if (lvl > 7 || (n_a << (7 - lvl)) != n_b) return ORC_INPUT_VALIDATION;
	uint32_t L = 1u << (7 - lvl);
	u128 acc = 0;
	for (uint64_t i = 0; i < n_a; i++)
		for (uint32_t j = 0; j < L; j++) acc ^= tower_mul_subfield(b[i * L + j], limb(a[i], lvl, j), lvl);
	*out = acc;
	return ORC_OK;
}

static int is_pow2(uint64_t x) { return x && !(x & (x - 1)); }
static uint32_t ilog2(uint64_t x) { uint32_t r = 0; while (x >>= 1) r++; return r; }

/* compute/src/cpu/layer.rs:238-258, 574-621   fold_left: out[i] = sum_j vec[j] * evals[j*rows + i] */
int orc_fold_left(const u128u *mat, uint64_t n_mat, uint32_t lvl, const u128u *vec, uint64_t n_vec,
				  u128u *out, uint64_t n_out) {
	tower_init();
	if (lvl > 7 || n_mat == 0 || n_vec == 0) return ORC_INPUT_VALIDATION;
	uint32_t L = 1u << (7 - lvl);
	uint32_t log_evals = ilog2(n_mat) + 7 - lvl;
	uint32_t log_q = ilog2(n_vec);
	if (log_q > log_evals) return ORC_INPUT_VALIDATION;
	uint64_t cols = (uint64_t)1 << log_q, rows = (uint64_t)1 << (log_evals - log_q);
	if (n_mat * L != cols * rows || n_vec != cols || n_out != rows) return ORC_INPUT_VALIDATION;
	for (uint64_t i = 0; i < rows; i++) {
		u128 acc = 0;
		for (uint64_t j = 0; j < cols; j++) {
			uint64_t e = j * rows + i;
			acc ^= tower_mul_subfield(vec[j], limb(mat[e / L], lvl, (uint32_t)(e % L)), lvl);
		}
		out[i] = acc;
	}
	return ORC_OK;
}

/* compute/src/cpu/layer.rs:260-280, 628-675   fold_right: out[i] = sum_j vec[j] * evals[i*rows_q + j] */
int orc_fold_right(const u128u *mat, uint64_t n_mat, uint32_t lvl, const u128u *vec, uint64_t n_vec,
				   u128u *out, uint64_t n_out) {
	tower_init();
	if (lvl > 7 || n_mat == 0 || n_vec == 0) return ORC_INPUT_VALIDATION;
	uint32_t L = 1u << (7 - lvl);
	uint32_t log_evals = ilog2(n_mat) + 7 - lvl;
	uint32_t log_q = ilog2(n_vec);
	if (log_q > log_evals) return ORC_INPUT_VALIDATION;
	uint64_t rows_q = (uint64_t)1 << log_q, cols = (uint64_t)1 << (log_evals - log_q);
	if (n_mat * L != cols * rows_q || n_vec != rows_q || n_out != cols) return ORC_INPUT_VALIDATION;
	for (uint64_t i = 0; i < cols; i++) {
		u128 acc = 0;
		for (uint64_t j = 0; j < rows_q; j++) {
			uint64_t e = i * rows_q + j;
			acc ^= tower_mul_subfield(vec[j], limb(mat[e / L], lvl, (uint32_t)(e % L)), lvl);
		}
		out[i] = acc;
	}
	return ORC_OK;
}

/* ---------------------------------------------------------------------------------------------
 * ArithCircuit evaluation -- math/src/arith_expr.rs:200-206 (steps), :367-383 (evaluate)
 * step encoding shared with include/binius_b200.h (b200_expr_step)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
	uint32_t op; /* 0 Add(l,r) 1 Mul(l,r) 2 Pow(l, r=exp) 3 Const(c) 4 Var(l) */
	uint32_t l;
	uint64_t r;
	uint64_t c_lo, c_hi;
} orc_expr_step;

static u128 b128_pow(u128 x, uint64_t e) {
	u128 r = 1;
	while (e) {
		if (e & 1) r = b128_mul(r, x);
		x = b128_mul(x, x);
		e >>= 1;
	}
	return r;
}

static u128 expr_eval(const orc_expr_step *steps, uint32_t n_steps, const u128 *query, u128 *tmp) {
	if (n_steps == 0) return 0;
	for (uint32_t s = 0; s < n_steps; s++) {
		const orc_expr_step *st = &steps[s];
		switch (st->op) {
		case 0: tmp[s] = tmp[st->l] ^ tmp[st->r]; break;
		case 1: tmp[s] = b128_mul(tmp[st->l], tmp[st->r]); break;
		case 2: tmp[s] = b128_pow(tmp[st->l], st->r); break;
		case 3: tmp[s] = ((u128)st->c_hi << 64) | st->c_lo; break;
		default: tmp[s] = query[st->l]; break;
		}
	}
	return tmp[n_steps - 1];
}

/* compute/src/cpu/layer.rs:410-435   compute_composite: out[i] = expr(inputs[.][i]) */
int orc_compute_composite(const u128u *const *inputs, uint32_t n_rows, uint64_t row_len, u128u *out,
						  uint64_t n_out, const orc_expr_step *steps, uint32_t n_steps, uint32_t n_vars) {
	tower_init();
	if (row_len != n_out || n_vars != n_rows) return ORC_INPUT_VALIDATION;
	u128 *tmp = malloc(sizeof(u128) * (n_steps + 1)), *q = malloc(sizeof(u128) * (n_rows + 1));
	for (uint64_t i = 0; i < n_out; i++) {
		for (uint32_t j = 0; j < n_rows; j++) q[j] = inputs[j][i];
		out[i] = expr_eval(steps, n_steps, q, tmp);
	}
	free(tmp); free(q);
	return ORC_OK;
}

/* compute/src/cpu/layer.rs:499-514   KernelExecutor::sum_composition_evals:
 *   acc += batch_coeff * sum_i expr(inputs[.][i]) */
int orc_sum_composition_evals(const u128u *const *inputs, uint32_t n_rows, uint64_t row_len,
							  const orc_expr_step *steps, uint32_t n_steps, const u128u *batch_coeff,
							  u128u *acc) {
	tower_init();
	u128 *tmp = malloc(sizeof(u128) * (n_steps + 1)), *q = malloc(sizeof(u128) * (n_rows + 1));
	u128 s = 0;
	for (uint64_t i = 0; i < row_len; i++) {
		for (uint32_t j = 0; j < n_rows; j++) q[j] = inputs[j][i];
		s ^= expr_eval(steps, n_steps, q, tmp);
	}
	*acc ^= b128_mul(s, *batch_coeff);
	free(tmp); free(q);
	return ORC_OK;
}

/* compute/src/cpu/layer.rs:437-484   pairwise_product_reduce; outs[r] has n >> (r+1) elements */
int orc_pairwise_product_reduce(const u128u *input, uint64_t n, u128u *const *outs, const uint64_t *out_lens,
								uint32_t n_outs) {
	tower_init();
	if (!is_pow2(n) || n < 2) return ORC_INPUT_VALIDATION;
	uint32_t log_n = ilog2(n);
	if (n_outs != log_n) return ORC_INPUT_VALIDATION;
	for (uint32_t r = 0; r < n_outs; r++)
		if (out_lens[r] != ((uint64_t)1 << (log_n - r - 1))) return ORC_INPUT_VALIDATION;
	const u128u *src = input;
	for (uint32_t r = 0; r < n_outs; r++) {
		for (uint64_t i = 0; i < out_lens[r]; i++) outs[r][i] = b128_mul(src[2 * i], src[2 * i + 1]);
		src = outs[r];
	}
	return ORC_OK;
}

/* ---------------------------------------------------------------------------------------------
 * core/src/protocols/sumcheck/v3/bivariate_product.rs:303-408   calculate_round_evals
 *   multilins[t] has 2^n elements; compositions c = (ia[c], ib[c]); batch coeff powers alpha^c
 *   out[0] = sum_c alpha^c sum_i hi_a[i]*hi_b[i]              (evaluation at 1)
 *   out[1] = sum_c alpha^c sum_i (lo_a+hi_a)[i]*(lo_b+hi_b)[i] (evaluation at infinity)
 * ------------------------------------------------------------------------------------------- */
int orc_bivariate_round_evals(const u128u *const *multilins, uint32_t n_multilins, uint32_t n_vars,
							  const uint32_t *ia, const uint32_t *ib, uint32_t n_comp,
							  const u128u *batch_coeff, u128u *out2) {
	tower_init();
	if (n_vars == 0) return ORC_INPUT_VALIDATION;
	uint64_t half = (uint64_t)1 << (n_vars - 1);
	u128 y1 = 0, yinf = 0, pw = 1;
	for (uint32_t c = 0; c < n_comp; c++) {
		if (ia[c] >= n_multilins || ib[c] >= n_multilins) return ORC_INPUT_VALIDATION;
		const u128u *a = multilins[ia[c]], *b = multilins[ib[c]];
		u128 s1 = 0, sinf = 0;
		for (uint64_t i = 0; i < half; i++) {
			s1 ^= b128_mul(a[half + i], b[half + i]);
			sinf ^= b128_mul(a[i] ^ a[half + i], b[i] ^ b[half + i]);
		}
		y1 ^= b128_mul(s1, pw);
		yinf ^= b128_mul(sinf, pw);
		pw = b128_mul(pw, *batch_coeff);
	}
	out2[0] = y1;
	out2[1] = yinf;
	return ORC_OK;
}

/* ---------------------------------------------------------------------------------------------
 * Additive NTT.
 *   ntt/src/twiddle.rs:244-313     precompute_subspace_evals (normalised subspace polynomial evals)
 *   ntt/src/twiddle.rs:141-168     OnTheFlyTwiddleAccess::get = subset_sum over the low log_n bits
 *   math/src/binary_subspace.rs:33-38  basis beta_j = 1 << j
 *   ntt/src/tests/reference.rs:68-160  forward/inverse_transform_simple
 *   ntt/src/additive_ntt.rs:8-27   batched layout: index = x | y << log_x | z << (log_x+log_y)
 *
 * The twiddle field is T_kt (kt = 3,4,5,6,7: B8..B128).  s_evals is returned flattened:
 * row r (r = 0..d-1) has d-1-r entries at offset r*(d-1) (padded rows; simple & small).
 * ------------------------------------------------------------------------------------------- */
int orc_ntt_s_evals(uint32_t kt, uint32_t d, u128u *s /* d*(d-1) */) {
	tower_init();
	if (d == 0 || d > (1u << kt)) return ORC_INPUT_VALIDATION;
	uint32_t W = d - 1;
	u128 *norm = malloc(sizeof(u128) * d);
	memset(s, 0, sizeof(u128) * d * (W ? W : 1));
	norm[0] = 1;
	for (uint32_t j = 0; j + 1 < d; j++) s[j] = (u128)1 << (j + 1);
	for (uint32_t r = 1; r < d; r++) {
		u128 np = norm[r - 1];
		const u128u *prev = &s[(r - 1) * W];
		/* subspace_map(e, c) = e^2 + c*e   (twiddle.rs:311-313) */
		norm[r] = tower_square(prev[0], kt) ^ tower_mul(np, prev[0], kt);
		/* prev row has d-r entries; row r = map over prev.skip(1) */
		for (uint32_t j = 1; j < d - r; j++)
			s[r * W + (j - 1)] = tower_square(prev[j], kt) ^ tower_mul(np, prev[j], kt);
	}
	for (uint32_t r = 0; r < d; r++) {
		u128 inv = tower_invert(norm[r], kt);
		for (uint32_t j = 0; j < d - 1 - r; j++) s[r * W + j] = tower_mul(s[r * W + j], inv, kt);
	}
	free(norm);
	return ORC_OK;
}

static inline u128 twiddle_get(const u128u *s, uint32_t d, uint32_t r, uint64_t idx) {
	uint32_t W = d - 1, log_n = d - 1 - r;
	u128 t = 0;
	for (uint32_t b = 0; b < log_n; b++)
		if ((idx >> b) & 1) t ^= s[r * W + b];
	return t;
}

/* AdditiveNTT::get_subspace_eval(i, j) = s_evals[d - i].get(j)   (single_threaded.rs:91-93) */
int orc_ntt_get_subspace_eval(const u128u *s, uint32_t d, uint32_t i, uint64_t j, u128u *out) {
	if (i > d || i == 0) return ORC_INPUT_VALIDATION;
	*out = twiddle_get(s, d, d - i, j);
	return ORC_OK;
}

static inline u128 load_elem(const uint8_t *p, uint32_t kd, uint64_t idx) {
	switch (kd) {
	case 3: return p[idx];
	case 4: return ((const uint16_t *)p)[idx];
	case 5: return ((const uint32_t *)p)[idx];
	case 6: return ((const uint64_t *)p)[idx];
	default: return ((const u128u *)p)[idx];
	}
}
static inline void store_elem(uint8_t *p, uint32_t kd, uint64_t idx, u128 v) {
	switch (kd) {
	case 3: p[idx] = (uint8_t)v; break;
	case 4: ((uint16_t *)p)[idx] = (uint16_t)v; break;
	case 5: ((uint32_t *)p)[idx] = (uint32_t)v; break;
	case 6: ((uint64_t *)p)[idx] = (uint64_t)v; break;
	default: ((u128u *)p)[idx] = v; break;
	}
}
/* a in T_kd times s in T_kt (kt <= kd): limb-wise */
static inline u128 mul_ext(u128 a, uint32_t kd, u128 s, uint32_t kt) {
	if (kd == kt) return tower_mul(a, s, kt);
	uint32_t w = 1u << kt, n = 1u << (kd - kt);
	u128 m = (((u128)1) << w) - 1, r = 0;
	for (uint32_t j = 0; j < n; j++) r |= tower_mul((a >> (j * w)) & m, s, kt) << (j * w);
	return r;
}

/* error codes mirror ntt/src/error.rs through include/binius_b200.h: 0 ok, else validation class */
#define ORC_NTT_POW2 11
#define ORC_NTT_SKIP 12
#define ORC_NTT_BATCH 13
#define ORC_NTT_COSET 14
#define ORC_NTT_DOMAIN 15

/* single_threaded.rs:364-406 check_batch_transform_inputs_and_params (with WIDTH = 1 scalars) */
static int ntt_check(uint32_t d, uint64_t n_elems, uint32_t log_x, uint32_t log_y, uint32_t log_z,
					 uint64_t coset, uint32_t coset_bits, uint32_t skip_rounds) {
	if (!is_pow2(n_elems)) return ORC_NTT_POW2;
	if (skip_rounds > log_y) return ORC_NTT_SKIP;
	uint64_t full_y = n_elems >> (log_x + log_z);
	if ((((uint64_t)1 << log_y) != full_y && n_elems > 2) || (((uint64_t)1 << log_y) > full_y)) return ORC_NTT_BATCH;
	if (coset >= ((uint64_t)1 << coset_bits)) return ORC_NTT_COSET;
	if (log_y + coset_bits > d) return ORC_NTT_DOMAIN;
	return ORC_OK;
}

int orc_ntt_transform(int inverse, const u128u *s, uint32_t kt, uint32_t d, uint8_t *data, uint32_t kd,
					  uint64_t n_elems, uint32_t log_x, uint32_t log_y, uint32_t log_z, uint64_t coset,
					  uint32_t coset_bits, uint32_t skip_rounds) {
	tower_init();
	if (kd < kt || kd > 7) return ORC_INPUT_VALIDATION;
	int rc = ntt_check(d, n_elems, log_x, log_y, log_z, coset, coset_bits, skip_rounds);
	if (rc) return rc;
	uint32_t row0 = d - (log_y + coset_bits);
	uint32_t n_layers = log_y - skip_rounds;
	for (uint64_t z = 0; z < ((uint64_t)1 << log_z); z++)
		for (uint64_t x = 0; x < ((uint64_t)1 << log_x); x++) {
			uint64_t base = x | z << (log_x + log_y);
			for (uint32_t li = 0; li < n_layers; li++) {
				uint32_t i = inverse ? li : (n_layers - 1 - li);
				for (uint64_t j = 0; j < ((uint64_t)1 << (log_y - 1 - i)); j++) {
					u128 t = twiddle_get(s, d, row0 + i, coset << (log_y - 1 - i) | j);
					for (uint64_t k = 0; k < ((uint64_t)1 << i); k++) {
						uint64_t i0 = j << (i + 1) | k, i1 = i0 | (uint64_t)1 << i;
						uint64_t p0 = base + (i0 << log_x), p1 = base + (i1 << log_x);
						u128 u = load_elem(data, kd, p0), v = load_elem(data, kd, p1);
						if (!inverse) { u ^= mul_ext(v, kd, t, kt); v ^= u; }
						else { v ^= u; u ^= mul_ext(v, kd, t, kt); }
						store_elem(data, kd, p0, u);
						store_elem(data, kd, p1, v);
					}
				}
			}
		}
	return ORC_OK;
}

/* ---------------------------------------------------------------------------------------------
 * compute/src/cpu/layer.rs:304-391   fri_fold (F = B128, FSub = T_kt twiddles from an NTT of dim d)
 * ------------------------------------------------------------------------------------------- */
static inline u128 lerp(u128 a, u128 b, u128 z) { return a ^ b128_mul(a ^ b, z); }

int orc_fri_fold(const u128u *s, uint32_t kt, uint32_t d, uint32_t log_len, uint32_t log_batch,
				 const u128u *challenges, uint32_t n_ch, const u128u *in, uint64_t n_in, u128u *out,
				 uint64_t n_out) {
	tower_init();
	if (n_in != ((uint64_t)1 << (log_len + log_batch))) return ORC_INPUT_VALIDATION;
	if (n_ch < log_batch) return ORC_INPUT_VALIDATION;
	if (n_ch > log_batch + log_len) return ORC_INPUT_VALIDATION;
	if (n_out != ((uint64_t)1 << (log_len - (n_ch - log_batch)))) return ORC_INPUT_VALIDATION;
	uint32_t eta = n_ch - log_batch;
	uint64_t chunk = (uint64_t)1 << n_ch;
	u128 *v = malloc(sizeof(u128) * chunk);
	for (uint64_t c = 0; c < n_out; c++) {
		for (uint64_t i = 0; i < chunk; i++) v[i] = in[c * chunk + i];
		uint64_t cur = chunk;
		for (uint32_t r = 0; r < log_batch; r++) {
			cur >>= 1;
			for (uint64_t o = 0; o < cur; o++) v[o] = lerp(v[2 * o], v[2 * o + 1], challenges[r]);
		}
		uint32_t L = log_len, sz = eta;
		for (uint32_t r = 0; r < eta; r++) {
			for (uint64_t o = 0; o < ((uint64_t)1 << (sz - 1)); o++) {
				u128 t = twiddle_get(s, d, d - L, (c << (sz - 1)) | o);
				u128 u = v[2 * o], w = v[2 * o + 1];
				w ^= u;
				u ^= mul_ext(w, 7, t, kt);
				v[o] = lerp(u, w, challenges[log_batch + r]);
			}
			L--; sz--;
		}
		out[c] = v[0];
	}
	free(v);
	return ORC_OK;
}

/* KernelExecutor::add / add_assign  (compute/src/cpu/layer.rs:516-548) */
int orc_add(const u128u *a, const u128u *b, u128u *dst, uint64_t n) {
	for (uint64_t i = 0; i < n; i++) dst[i] = a[i] ^ b[i];
	return ORC_OK;
}

/* scalar helpers exported for tests */
void orc_mul(const u128u *a, const u128u *b, uint32_t k, u128u *out) { tower_init(); *out = tower_mul(*a, *b, k); }
void orc_mul_slow(const u128u *a, const u128u *b, uint32_t k, u128u *out) { *out = tower_mul_slow(*a, *b, k); }
void orc_square(const u128u *a, uint32_t k, u128u *out) { tower_init(); *out = tower_square(*a, k); }
void orc_invert(const u128u *a, uint32_t k, u128u *out) { tower_init(); *out = tower_invert(*a, k); }
void orc_mul_alpha(const u128u *a, uint32_t k, u128u *out) { tower_init(); *out = tower_mul_alpha(*a, k); }
void orc_mul_subfield(const u128u *a, const u128u *s, uint32_t k, u128u *out) { tower_init(); *out = tower_mul_subfield(*a, *s, k); }
void orc_mul_vec(const u128u *a, const u128u *b, u128u *out, uint64_t n) {
	tower_init();
	for (uint64_t i = 0; i < n; i++) out[i] = b128_mul(a[i], b[i]);
}
