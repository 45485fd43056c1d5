// univariate.cuh -- the zerocheck univariate-skip round.
//
// Reference: zerocheck_univariate_evals, core/src/protocols/sumcheck/prove/univariate.rs:235-500, with
// ntt_extrapolate (:642-678), spread_product (:503-563) and extrapolate_round_evals (:565-640).
//
//   R[c][i] = sum_{s < 2^(n-k)} eq[s] * C_c( P_0(s, x_i), ..., P_{m-1}(s, x_i) ),      x_i = B8(2^k + i)
//   P_j(s, x) = extrapolation of M_j[s*2^k .. (s+1)*2^k) from the B8 points 0..2^k-1 to x
//
// The reference extrapolates with an inverse + coset-forward additive NTT over B8; any exact evaluation of
// the same polynomial is bit-identical, and this kernel uses the Lagrange form with the B8 coefficients
// L_t(x_i) staged in shared memory: a B1 column needs no multiplication at all (the extrapolated value is
// the XOR of the coefficients selected by the sub-cube's bits, four bits per nibble-table lookup).  All
// composition arithmetic runs in the base field FBase = smallest tower level >= B8 holding every column and
// constant (B8 products are one shared-memory table lookup), as the reference's PackedSubfield<P, FBase>
// evaluation does; only the final eq[s] * value product touches B128 (limb-wise, binary_field.rs:363-393).
//
// Thread mapping: a CTA owns 2^k consecutive evaluation points (blockIdx.y) and walks sub-cubes
// (blockIdx.x, grid-stride); thread = (point, sub-cube lane).  Per-thread accumulators, one per
// composition, live in local memory; they are combined across sub-cube lanes in shared memory and across
// CTAs by XOR atomics on the [composition][point] table.
#pragma once
#include "field.cuh"
#include "kernels.cuh"

namespace b200 {
namespace uni {

constexpr uint32_t THREADS = 256;
constexpr uint32_t MAX_COMP = 128;  // compositions per call (accumulators per thread)
constexpr uint32_t MAX_MLS = 256;   // multilinears per call

struct Args {
	const uint4 *const *mls;   // device [m]: packed sub-field multilinears
	const uint32_t *levels;    // device [m]: their tower levels
	const DevExpr *comps;      // device [n_comp]
	const uint32_t *comp_pts;  // device [n_comp]: (deg_c - 1) << skip evaluation points of composition c
	const uint8_t *lag;        // device [n_pts][2^skip]: L_t(x_i) in B8
	const uint4 *eq;           // 2^(n_vars - skip) B128
	uint4 *out;                // [n_comp][n_out] accumulators (zeroed)
	uint64_t n_sub;
	uint32_t m, n_comp, skip, n_pts, n_out;
	uint32_t off_nl, off_red;  // shared-memory offsets of the nibble table and the reduction buffer
};

// base-field value types
template <uint32_t LVL> struct Fb;
template <> struct Fb<3> {
	typedef uint32_t V;
	static __device__ __forceinline__ V from_u32(uint32_t x) { return x; }
	static __device__ __forceinline__ V from128(uint4 a) { return a.x & 0xffu; }
	static __device__ __forceinline__ uint4 to128(V a) { return make_uint4(a, 0, 0, 0); }
	static __device__ __forceinline__ V mul(const FieldTables &T, V a, V b) { return f_mul8(T, a, b); }
	static __device__ __forceinline__ V mulc(const FieldTables &T, V a, uint32_t c) { return f_mul8(T, c, a); }  // uniform operand first (field.cuh)
};
template <> struct Fb<4> {
	typedef uint32_t V;
	static __device__ __forceinline__ V from_u32(uint32_t x) { return x; }
	static __device__ __forceinline__ V from128(uint4 a) { return a.x & 0xffffu; }
	static __device__ __forceinline__ uint4 to128(V a) { return make_uint4(a, 0, 0, 0); }
	static __device__ __forceinline__ V mul(const FieldTables &T, V a, V b) { return f_mul16(T, a, b); }
	static __device__ __forceinline__ V mulc(const FieldTables &T, V a, uint32_t c) { return f_mul8(T, c, a & 0xff) | (f_mul8(T, c, a >> 8) << 8); }
};
template <> struct Fb<5> {
	typedef uint32_t V;
	static __device__ __forceinline__ V from_u32(uint32_t x) { return x; }
	static __device__ __forceinline__ V from128(uint4 a) { return a.x; }
	static __device__ __forceinline__ uint4 to128(V a) { return make_uint4(a, 0, 0, 0); }
	static __device__ __forceinline__ V mul(const FieldTables &T, V a, V b) { return f_mul32(T, a, b); }
	static __device__ __forceinline__ V mulc(const FieldTables &T, V a, uint32_t c) { return f_mul8x4(T, a, c); }
};
template <> struct Fb<7> {
	typedef uint4 V;
	static __device__ __forceinline__ V from_u32(uint32_t x) { return make_uint4(x, 0, 0, 0); }
	static __device__ __forceinline__ V from128(uint4 a) { return a; }
	static __device__ __forceinline__ uint4 to128(V a) { return a; }
	static __device__ __forceinline__ V mul(const FieldTables &T, V a, V b) { return f_mul128(T, a, b); }
	static __device__ __forceinline__ V mulc(const FieldTables &T, V a, uint32_t c) { return f_mul128_sub(T, a, make_uint4(c, 0, 0, 0), 3); }
};

// scalar `idx` of a packed sub-field multilinear (2^(7-lvl) scalars per B128 word, low limb first)
template <uint32_t LVL>
__device__ __forceinline__ typename Fb<LVL>::V load_scalar(const uint4 *ml, uint32_t lvl, uint64_t idx) {
	typedef Fb<LVL> F;
	switch (lvl) {
	case 0: return F::from_u32((__ldg(reinterpret_cast<const uint32_t *>(ml) + (idx >> 5)) >> (idx & 31)) & 1u);
	case 3: return F::from_u32(__ldg(reinterpret_cast<const uint8_t *>(ml) + idx));
	case 4: return F::from_u32(__ldg(reinterpret_cast<const uint16_t *>(ml) + idx));
	case 5: return F::from_u32(__ldg(reinterpret_cast<const uint32_t *>(ml) + idx));
	default:
		if constexpr (LVL == 7) {
			if (lvl == 6) {
				uint2 v = __ldg(reinterpret_cast<const uint2 *>(ml) + idx);
				return make_uint4(v.x, v.y, 0, 0);
			}
			return __ldg(ml + idx);
		} else {
			return F::from_u32(0);
		}
	}
}

// ArithCircuit over the base field (math/src/arith_expr.rs:367-383)
template <uint32_t LVL>
__device__ __forceinline__ typename Fb<LVL>::V expr_eval_fb(const FieldTables &T, const DevExpr &E, const typename Fb<LVL>::V *q) {
	typedef Fb<LVL> F;
	typename F::V tmp[MAX_EXPR_STEPS];
	for (uint32_t s = 0; s < E.n_steps; s++) {
		const b200_expr_step st = E.steps[s];
		typename F::V v;
		switch (st.op) {
		case 0: v = tmp[st.l] ^ tmp[st.r]; break;
		case 1: v = F::mul(T, tmp[st.l], tmp[st.r]); break;
		case 2: {
			typename F::V x = tmp[st.l];
			uint64_t e = st.r;
			v = F::from_u32(1);
			while (e) {
				if (e & 1) v = F::mul(T, v, x);
				e >>= 1;
				if (e) x = F::mul(T, x, x);
			}
			break;
		}
		case 3: v = F::from128(make_uint4((uint32_t)st.c_lo, (uint32_t)(st.c_lo >> 32), (uint32_t)st.c_hi, (uint32_t)(st.c_hi >> 32))); break;
		default: v = q[st.l]; break;
		}
		tmp[s] = v;
	}
	return E.n_steps ? tmp[E.n_steps - 1] : F::from_u32(0);
}

template <uint32_t LVL>
__global__ void __launch_bounds__(THREADS) k_uni_evals(const uint8_t *__restrict__ g_tables, const Args A) {
	typedef Fb<LVL> F;
	typedef typename F::V V;
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	const uint32_t K = 1u << A.skip, tid = threadIdx.x;
	uint8_t *lagS = smem + FIELD_TABLE_BYTES;  // [t][point]: L_t(x_point)
	uint8_t *NL = smem + A.off_nl;             // [nibble][pattern][point]: XOR of the coefficients the pattern selects
	uint4 *red = reinterpret_cast<uint4 *>(smem + A.off_red);
	const uint32_t p0 = blockIdx.y * K;
	for (uint32_t idx = tid; idx < K * K; idx += THREADS) {
		uint32_t t = idx >> A.skip, il = idx & (K - 1);
		lagS[idx] = A.lag[(uint64_t)(p0 + il) * K + t];
	}
	__syncthreads();
	const bool use_nl = K >= 4;
	if (use_nl) {
		for (uint32_t idx = tid; idx < 4 * K * K; idx += THREADS) {
			uint32_t il = idx & (K - 1), e = idx >> A.skip, pat = e & 15, nib = e >> 4;
			uint32_t v = 0;
#pragma unroll
			for (uint32_t b = 0; b < 4; b++)
				if (pat >> b & 1) v ^= lagS[((4 * nib + b) << A.skip) + il];
			NL[idx] = (uint8_t)v;
		}
		__syncthreads();
	}
	const uint32_t il = tid & (K - 1), sl = tid >> A.skip, SL = THREADS >> A.skip;
	const uint32_t i = p0 + il;

	uint4 acc[MAX_COMP];
	V q[MAX_MLS];
	for (uint32_t c = 0; c < A.n_comp; c++) acc[c] = u4_zero();

	for (uint64_t s = (uint64_t)blockIdx.x * SL + sl; s < A.n_sub; s += (uint64_t)gridDim.x * SL) {
		const uint64_t base = s << A.skip;
		for (uint32_t j = 0; j < A.m; j++) {
			const uint4 *ml = A.mls[j];
			const uint32_t lvl = A.levels[j];
			V v = F::from_u32(0);
			if (lvl == 0 && use_nl) {
				const uint32_t *w = reinterpret_cast<const uint32_t *>(ml) + (base >> 5);
				if (K >= 32) {
					for (uint32_t ww = 0; ww < (K >> 5); ww++) {
						uint32_t bits = __ldg(w + ww);
						uint32_t x = 0;
#pragma unroll
						for (uint32_t n = 0; n < 8; n++) x ^= NL[((((ww << 3) + n) << 4) + ((bits >> (4 * n)) & 15u)) * K + il];
						v = v ^ F::from_u32(x);
					}
				} else {
					uint32_t bits = __ldg(w) >> (base & 31);
					uint32_t x = 0;
					for (uint32_t n = 0; n < (K >> 2); n++) x ^= NL[((n << 4) + ((bits >> (4 * n)) & 15u)) * K + il];
					v = F::from_u32(x);
				}
			} else if (lvl == 0) {
				for (uint32_t t = 0; t < K; t++) {
					uint32_t bit = (__ldg(reinterpret_cast<const uint32_t *>(ml) + ((base + t) >> 5)) >> ((base + t) & 31)) & 1u;
					if (bit) v = v ^ F::from_u32(lagS[(t << A.skip) + il]);
				}
			} else {
				for (uint32_t t = 0; t < K; t++) v = v ^ F::mulc(T, load_scalar<LVL>(ml, lvl, base + t), lagS[(t << A.skip) + il]);
			}
			q[j] = v;
		}
		const uint4 e = __ldg(A.eq + s);
		for (uint32_t c = 0; c < A.n_comp; c++) {
			if (i >= A.comp_pts[c]) continue;
			V val = expr_eval_fb<LVL>(T, A.comps[c], q);
			acc[c] ^= f_mul128_sub(T, e, F::to128(val), LVL);
		}
	}
	for (uint32_t c = 0; c < A.n_comp; c++) {
		red[tid] = acc[c];
		__syncthreads();
		if (sl == 0 && i < A.comp_pts[c]) {
			uint4 v = red[il];
			for (uint32_t l = 1; l < SL; l++) v ^= red[(l << A.skip) + il];
			atomic_xor_u4(A.out + (uint64_t)c * A.n_out + i, v);
		}
		__syncthreads();
	}
}

// ---- B8 fast path ------------------------------------------------------------------------------------
// FBase = B8 (every column B1 or B8, constants in B8) and every composition a sum of monomials of degree
// <= 2 (eqind_plan.hpp) -- the keccak / u32 gadget shape.  Work is organised in batches of SUBS sub-cubes
// per CTA so that nothing but the final accumulators leaves shared memory:
//   phase A: thread = (quad of 4 points, sub-cube, column lane): extrapolates 4 points at once -- the nibble
//            table holds the four B8 coefficients of a point quad in one 32-bit word, so a 64-bit B1
//            sub-cube costs 16 LDS.32 + 16 XOR for 4 points -- and stores them to qS[column][sub-cube][point].
//   phase B: thread = (point, composition lane): evaluates its compositions from the monomial list on qS
//            bytes (one mul8 table lookup per product) and multiplies by eq[s] through two 16-entry B128
//            tables eq[s]*n and eq[s]*(n<<4) built once per sub-cube (x -> eq[s]*x is GF(2)-linear), i.e.
//            8 conflict-free LDS.32 instead of 16 byte products; sums over the batch stay in registers, the per-thread
//            accumulator (local memory) is touched once per composition and batch.
constexpr uint32_t B8_THREADS = 512;
constexpr uint32_t SUBS = 8;
constexpr uint32_t MONO_NONE = 511;
constexpr uint32_t CTAB = 5;  // words per composition: first monomial, #quadratic (coef 1), #linear (coef 1), #general, points
struct B8Args {
	const uint4 *const *mls;   // device [m]
	const uint32_t *levels;    // device [m] (0 or 3)
	const uint2 *mono;         // device: per monomial {a * SUBS * K, b * SUBS * K} (byte offsets into qS) for the coef-1
							   // quadratic / linear runs, {a | b << 9 | coef << 18, 0} for the general run
	const uint32_t *comp_tab;  // device [n_comp][CTAB]
	const uint8_t *lag;        // device [n_pts][K]
	const uint4 *eq;
	uint4 *out;
	uint64_t n_sub;
	uint32_t m, n_comp, n_mono, n_out;
	uint32_t off_nl, off_q, off_es, off_mono, off_ctab, off_cols, off_bits;  // shared-memory layout
	// PREP (challenge-independent half of the round, see k_uni_finish): the B8 value of every (batch, composition, point) goes to
	// store[batch][composition][point][sub-cube of the batch] instead of being weighted by eq
	uint8_t *store;       // already offset to the first composition of this launch and the first batch of this chunk
	uint64_t rec_bytes;   // one batch of ALL compositions: n_comp_total * n_pts * SUBS
	uint32_t n_pts;
};

template <uint32_t SKIP, bool PREP = false>
__global__ void __launch_bounds__(B8_THREADS, 1) k_uni_b8(const uint8_t *__restrict__ g_tables, const B8Args A) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	constexpr uint32_t K = 1u << SKIP, PQ = K >> 2;
	const uint32_t tid = threadIdx.x;
	uint8_t *lagS = smem + FIELD_TABLE_BYTES;
	uint32_t *NLw = reinterpret_cast<uint32_t *>(smem + A.off_nl);  // [nibble][pattern][32 words = 128/K copies of the K points]
	uint8_t *qS = smem + A.off_q;                                   // [column][sub-cube][point]
	uint32_t *ESw = reinterpret_cast<uint32_t *>(smem + A.off_es);  // [word][sub-cube][32]
	uint2 *monoS = reinterpret_cast<uint2 *>(smem + A.off_mono);
	uint32_t *ctabS = reinterpret_cast<uint32_t *>(smem + A.off_ctab);
	const uint32_t **colP = reinterpret_cast<const uint32_t **>(smem + A.off_cols);
	uint32_t *colL = reinterpret_cast<uint32_t *>(smem + A.off_cols + 8 * A.m);
	uint32_t *bitsS = reinterpret_cast<uint32_t *>(smem + A.off_bits);  // [column][WPC]: the batch's B1 sub-cubes
	const uint32_t p0 = blockIdx.y * K;
	for (uint32_t idx = tid; idx < K * K; idx += B8_THREADS) lagS[idx] = A.lag[(uint64_t)(p0 + (idx & (K - 1))) * K + (idx >> SKIP)];
	for (uint32_t idx = tid; idx < A.n_mono; idx += B8_THREADS) monoS[idx] = A.mono[idx];
	for (uint32_t idx = tid; idx < CTAB * A.n_comp; idx += B8_THREADS) ctabS[idx] = A.comp_tab[idx];
	for (uint32_t idx = tid; idx < A.m; idx += B8_THREADS) {
		colP[idx] = reinterpret_cast<const uint32_t *>(A.mls[idx]);
		colL[idx] = A.levels[idx];
	}
	// stale bytes of qS are read (and multiplied by a zero eq table) for sub-cubes past the end: keep them defined
	for (uint32_t idx = tid; idx < (A.m * SUBS * K) / 4; idx += B8_THREADS) reinterpret_cast<uint32_t *>(qS)[idx] = 0;
	__syncthreads();
	// rows are 32 words: 128 / K copies of the K coefficient bytes side by side, so that the 32 / PQ different
	// (sub-cube, column) pairs a warp works on read their rows from disjoint banks (lane l reads word l)
	for (uint32_t idx = tid; idx < PQ * 16 * 32; idx += B8_THREADS) {
		uint32_t pq = (idx & 31) % PQ, e = idx >> 5, pat = e & 15, nib = e >> 4, word = 0;
#pragma unroll
		for (uint32_t p = 0; p < 4; p++) {
			uint32_t v = 0;
#pragma unroll
			for (uint32_t b = 0; b < 4; b++)
				if (pat >> b & 1) v ^= lagS[((4 * nib + b) << SKIP) + 4 * pq + p];
			word |= v << (8 * p);
		}
		NLw[idx] = word;
	}
	__syncthreads();
	// phase A coordinates
	constexpr uint32_t JL = B8_THREADS / (PQ * SUBS);
	const uint32_t a_pq = tid % PQ, a_r = tid / PQ, a_sb = a_r % SUBS, a_jl = a_r / SUBS;
	// phase B coordinates
	constexpr uint32_t G = B8_THREADS >> SKIP;
	const uint32_t il = tid & (K - 1), g = tid >> SKIP, i = p0 + il;
	uint4 accL[MAX_COMP / 4];
#pragma unroll
	for (uint32_t c = 0; c < MAX_COMP / 4; c++) accL[c] = u4_zero();

	// B1 column words of a batch (WPC consecutive 32-bit words per column) are fetched one batch ahead into
	// registers while phase B runs, and parked in shared memory for phase A
	constexpr uint32_t WPC = (SUBS * K) / 32, NPRE = 8;
	const uint32_t n_words = A.m * WPC;  // <= NPRE * B8_THREADS (host-checked)
	const uint64_t col_words = max((A.n_sub << SKIP) >> 5, (uint64_t)1);
	uint32_t pre[NPRE];
	auto prefetch = [&](uint64_t batch) {
#pragma unroll
		for (uint32_t r = 0; r < NPRE; r++) {
			const uint32_t idx = tid + r * B8_THREADS, j = idx / WPC, w = idx % WPC;
			const uint64_t wi = batch * WPC + w;
			pre[r] = (idx < n_words && colL[j] == 0 && wi < col_words) ? __ldg(colP[j] + wi) : 0u;
		}
	};
	auto park = [&]() {
#pragma unroll
		for (uint32_t r = 0; r < NPRE; r++) {
			const uint32_t idx = tid + r * B8_THREADS;
			if (idx < n_words) bitsS[idx] = pre[r];
		}
	};
	const uint64_t n_batches = (A.n_sub + SUBS - 1) / SUBS;
	if (blockIdx.x < n_batches) {
		prefetch(blockIdx.x);
		park();
	}
	__syncthreads();
	for (uint64_t bt = blockIdx.x; bt < n_batches; bt += gridDim.x) {
		const uint64_t s0 = bt * SUBS;
		if (!PREP && tid < SUBS * 32) {
			uint32_t sb = tid >> 5, e = tid & 31;
			uint4 eqv = s0 + sb < A.n_sub ? __ldg(A.eq + s0 + sb) : u4_zero();
			// four 32-bit planes [word][sub-cube][32]: the 16 entries of a nibble table sit in 16 distinct banks,
			// so the gathers of phase B are conflict-free whatever the values (an LDS.128 gather is not)
			const uint4 v = f_mul128_b8_uniform(T, eqv, e < 16 ? e : (e - 16) << 4);  // (eq[s] is uniform in the warp, the entry varies)
			ESw[tid] = v.x, ESw[SUBS * 32 + tid] = v.y, ESw[2 * SUBS * 32 + tid] = v.z, ESw[3 * SUBS * 32 + tid] = v.w;
		}
		const uint64_t s = s0 + a_sb;
		if (s < A.n_sub) {
			const uint64_t base = s << SKIP;
			const uint32_t *nl = NLw + (tid & 31);
			for (uint32_t j = a_jl; j < A.m; j += JL) {
				uint32_t x = 0;
				if (colL[j] == 0) {
					if constexpr (K >= 32) {
						const uint32_t *w = bitsS + j * WPC + a_sb * (K >> 5);
#pragma unroll
						for (uint32_t ww = 0; ww < (K >> 5); ww++) {
							const uint32_t bits = w[ww];
#pragma unroll
							for (uint32_t n = 0; n < 8; n++) x ^= nl[((((ww << 3) + n) << 4) + ((bits >> (4 * n)) & 15u)) << 5];
						}
					} else {
						const uint32_t bits = bitsS[j * WPC + ((a_sb * K) >> 5)] >> ((a_sb * K) & 31);
#pragma unroll
						for (uint32_t n = 0; n < (K >> 2); n++) x ^= nl[((n << 4) + ((bits >> (4 * n)) & 15u)) << 5];
					}
				} else {
					const uint8_t *col = reinterpret_cast<const uint8_t *>(colP[j]) + base;
					for (uint32_t t = 0; t < K; t++) {
						const uint32_t mv = (uint32_t)__ldg(col + t) << 8;
						const uint32_t lw = *reinterpret_cast<const uint32_t *>(lagS + (t << SKIP) + 4 * a_pq);
						x ^= (uint32_t)T.mul8[mv | (lw & 0xff)] | ((uint32_t)T.mul8[mv | ((lw >> 8) & 0xff)] << 8) |
							 ((uint32_t)T.mul8[mv | ((lw >> 16) & 0xff)] << 16) | ((uint32_t)T.mul8[mv | (lw >> 24)] << 24);
					}
				}
				reinterpret_cast<uint32_t *>(qS)[(j * SUBS + a_sb) * PQ + a_pq] = x;
			}
		}
		__syncthreads();
		const bool more = bt + gridDim.x < n_batches;
		if (more) prefetch(bt + gridDim.x);
		const uint8_t *qb = qS + il;
		for (uint32_t c = g, k = 0; c < A.n_comp; c += G, k++) {
			const uint32_t *ct = ctabS + CTAB * c;
			if (i >= ct[4]) continue;
			const uint2 *mp = monoS + ct[0];
			uint32_t val[SUBS];
#pragma unroll
			for (uint32_t sb = 0; sb < SUBS; sb++) val[sb] = 0;
			for (uint32_t t = 0; t < ct[1]; t++) {  // x_a * x_b
				const uint2 d = mp[t];
#pragma unroll
				for (uint32_t sb = 0; sb < SUBS; sb++) val[sb] ^= T.mul8[((uint32_t)qb[d.x + sb * K] << 8) | qb[d.y + sb * K]];
			}
			mp += ct[1];
			for (uint32_t t = 0; t < ct[2]; t++) {  // x_a
				const uint32_t d = mp[t].x;
#pragma unroll
				for (uint32_t sb = 0; sb < SUBS; sb++) val[sb] ^= qb[d + sb * K];
			}
			mp += ct[2];
			for (uint32_t t = 0; t < ct[3]; t++) {  // coef * (1 | x_a | x_a * x_b)
				const uint32_t d = mp[t].x, a = d & 511u, b = (d >> 9) & 511u, cf = d >> 18;
#pragma unroll
				for (uint32_t sb = 0; sb < SUBS; sb++) {
					uint32_t v = cf;
					if (a != MONO_NONE) {
						v = qb[(a * SUBS + sb) * K];
						if (b != MONO_NONE) v = T.mul8[(v << 8) | qb[(b * SUBS + sb) * K]];
						v = T.mul8[(v << 8) | cf];
					}
					val[sb] ^= v;
				}
			}
			if constexpr (PREP) {
				static_assert(SUBS == 8, "one uint2 per (batch, composition, point)");
				const uint2 w = make_uint2(val[0] | (val[1] << 8) | (val[2] << 16) | (val[3] << 24), val[4] | (val[5] << 8) | (val[6] << 16) | (val[7] << 24));
				*reinterpret_cast<uint2 *>(A.store + bt * A.rec_bytes + ((uint64_t)c * A.n_pts + i) * SUBS) = w;
			} else {
				uint4 acc = u4_zero();
#pragma unroll
				for (uint32_t sb = 0; sb < SUBS; sb++) {
					const uint32_t *lo = ESw + sb * 32 + (val[sb] & 15u), *hi = ESw + sb * 32 + 16 + (val[sb] >> 4);
					acc.x ^= lo[0] ^ hi[0];
					acc.y ^= lo[SUBS * 32] ^ hi[SUBS * 32];
					acc.z ^= lo[2 * SUBS * 32] ^ hi[2 * SUBS * 32];
					acc.w ^= lo[3 * SUBS * 32] ^ hi[3 * SUBS * 32];
				}
				accL[k] ^= acc;
			}
		}
		if (more) park();
		__syncthreads();
	}
	if constexpr (!PREP)
		for (uint32_t c = g, k = 0; c < A.n_comp; c += G, k++)
			if (i < ctabS[CTAB * c + 4]) atomic_xor_u4(A.out + (uint64_t)c * A.n_out + i, accL[k]);
}

// The challenge-dependent half of the B8 fast path.  k_uni_b8<SKIP, true> (which needs the witness and the constraints
// but no challenge, so it can run while the witness is still being uploaded / committed) has left the B8 value
// v[s][c][i] of every composition's non-linear part at every extrapolation point in `store`; what remains of the round is
//     out[c][i] ^= sum_s eq[s] * v[s][c][i]
// i.e. per batch of SUBS sub-cubes the two nibble tables eq[s]*n, eq[s]*(n<<4) per sub-cube (as in k_uni_b8) and 16
// conflict-free gathers of four 32-bit planes per (composition, point).  The records stream from HBM through a
// two-deep cp.async ring (one batch of a composition range = nc * n_pts * SUBS contiguous bytes).
// grid = (batches (grid-stride), composition ranges), block = B8_THREADS
struct FinArgs {
	const uint8_t *store;
	const uint32_t *comp_pts;  // device [n_comp]
	const uint4 *eq;
	uint4 *out;                // [n_comp][n_out] (XOR-accumulated)
	uint64_t n_sub, rec_bytes;
	uint32_t n_comp, n_pts, n_out, range;  // compositions per blockIdx.y
	uint32_t buf_bytes;                     // bytes of one ring slot (>= range * n_pts * SUBS, multiple of 16)
};
constexpr uint32_t FIN_ACC = 10;  // (composition, point) pairs per thread, accumulators in registers: range * n_pts <= FIN_ACC * B8_THREADS
// The 64 value bits of a batch (8 sub-cubes x 8 bits) are cut into 13 groups of 5 bits (the last one 4): a 32-entry table of
// 32-bit planes is one row of the 32 banks, so its gathers are conflict-free whatever the indices -- like the 16-entry
// nibble tables of k_uni_b8, with 13 instead of 16 gathers per plane.
constexpr uint32_t FIN_GROUPS = 13, FIN_ES_BYTES = 4 * FIN_GROUPS * 32 * 4;
__global__ void __launch_bounds__(B8_THREADS, 1) k_uni_finish(const uint8_t *__restrict__ g_tables, const FinArgs A) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	uint32_t *ESw = reinterpret_cast<uint32_t *>(smem + FIELD_TABLE_BYTES);  // [word][group][32]
	uint32_t *ptsS = ESw + FIN_ES_BYTES / 4;                                  // [range]
	uint8_t *buf = smem + FIELD_TABLE_BYTES + FIN_ES_BYTES + 4 * MAX_COMP;
	const uint32_t tid = threadIdx.x, c0 = blockIdx.y * A.range, nc = min(A.range, A.n_comp - c0), total = nc * A.n_pts;
	for (uint32_t idx = tid; idx < nc; idx += B8_THREADS) ptsS[idx] = A.comp_pts[c0 + idx];
	uint4 accL[FIN_ACC];
#pragma unroll
	for (uint32_t k = 0; k < FIN_ACC; k++) accL[k] = u4_zero();
	const uint64_t n_batches = (A.n_sub + SUBS - 1) / SUBS;
	const uint32_t slice = total * SUBS;  // multiple of 16: n_pts is a multiple of 4
	const uint8_t *src0 = A.store + (uint64_t)c0 * A.n_pts * SUBS;
	auto fetch = [&](uint64_t bt, uint32_t slot) {
		const uint8_t *src = src0 + bt * A.rec_bytes;
		const uint32_t dst = (uint32_t)__cvta_generic_to_shared(buf + slot * A.buf_bytes);
		for (uint32_t off = tid * 16; off < slice; off += B8_THREADS * 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + off), "l"(src + off) : "memory");
		asm volatile("cp.async.commit_group;" ::: "memory");
	};
	if (blockIdx.x < n_batches) fetch(blockIdx.x, 0);
	__syncthreads();
	uint32_t it = 0;
	for (uint64_t bt = blockIdx.x; bt < n_batches; bt += gridDim.x, it++) {
		const bool more = bt + gridDim.x < n_batches;
		if (more) fetch(bt + gridDim.x, (it + 1) & 1);
		if (tid < FIN_GROUPS * 32) {
			// entry e of group g = sum of eq[sub-cube of bit 5g+k] * 2^((5g+k) mod 8) over the set bits k of e: the group's bits lie
			// in at most two sub-cubes, so it is two B128 x B8 products
			const uint32_t g = tid >> 5, e = tid & 31, b0 = 5 * g, sb0 = b0 >> 3, o0 = b0 & 7;
			const uint32_t in0 = min(8u - o0, 5u);  // bits of the group that belong to the first sub-cube
			const uint64_t s0 = bt * SUBS + sb0;
			const uint4 eq0 = s0 < A.n_sub ? __ldg(A.eq + s0) : u4_zero();
			uint4 v = f_mul128_b8_uniform(T, eq0, (e & ((1u << in0) - 1)) << o0);  // (eq is uniform in the warp, the entry varies)
			if (in0 < 5 && sb0 + 1 < SUBS) {
				const uint4 eq1 = s0 + 1 < A.n_sub ? __ldg(A.eq + s0 + 1) : u4_zero();
				v ^= f_mul128_b8_uniform(T, eq1, e >> in0);
			}
			ESw[tid] = v.x, ESw[FIN_GROUPS * 32 + tid] = v.y, ESw[2 * FIN_GROUPS * 32 + tid] = v.z, ESw[3 * FIN_GROUPS * 32 + tid] = v.w;
		}
		if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
		else asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncthreads();
		const uint2 *rb = reinterpret_cast<const uint2 *>(buf + (it & 1) * A.buf_bytes);
#pragma unroll
		for (uint32_t k = 0; k < FIN_ACC; k++) {  // fully unrolled: accL stays in registers
			const uint32_t p = tid + k * B8_THREADS;
			if (p >= total) break;
			const uint32_t c = p / A.n_pts, i = p - c * A.n_pts;
			if (i >= ptsS[c]) continue;
			const uint2 w = rb[p];
			uint4 acc = u4_zero();
#pragma unroll
			for (uint32_t g = 0; g < FIN_GROUPS; g++) {
				const uint32_t b0 = 5 * g;
				const uint32_t ix = (b0 + 5 <= 32 ? w.x >> b0 : b0 >= 32 ? w.y >> (b0 - 32) : __funnelshift_r(w.x, w.y, b0)) & 31u;
				const uint32_t *e = ESw + g * 32 + ix;
				acc.x ^= e[0];
				acc.y ^= e[FIN_GROUPS * 32];
				acc.z ^= e[2 * FIN_GROUPS * 32];
				acc.w ^= e[3 * FIN_GROUPS * 32];
			}
			accL[k] ^= acc;
		}
		__syncthreads();
	}
#pragma unroll
	for (uint32_t k = 0; k < FIN_ACC; k++) {
		const uint32_t p = tid + k * B8_THREADS;
		if (p >= total) break;
		const uint32_t c = p / A.n_pts, i = p - c * A.n_pts;
		if (i < ptsS[c]) atomic_xor_u4(A.out + (uint64_t)(c0 + c) * A.n_out + i, accL[k]);
	}
}

// Linear monomials of the B8 fast path (skip = 7, B1 columns).  sum_s eq[s] * P_j(s, x_i) is linear in the column:
//     sum_s eq[s] * sum_t bit_j[s][t] * L_t(x_i)  =  sum_t L_t(x_i) * E_j[t],     E_j[t] = sum_s eq[s] * bit_j[s*128 + t],
// and E_j is evaluate_partial_high of the column by eq (an outer-product bit-GEMM on the tensor cores, roundevals_tc.cuh).
// So coefficient-1 linear monomials never enter k_uni_b8 (no extrapolation of columns that occur only linearly, fewer
// gathers per composition); this kernel adds  sum_t L_t(x_i) * (sum_{j in lin(c)} E_j[t])  to out[c][i].
// grid = n_comp, block = 128 (= 2^skip points per degree), dyn smem = FIELD_TABLE_BYTES
struct LinArgs {
	const uint32_t *lin_off;   // device [n_comp + 1]: range of composition c in lin_cols
	const uint32_t *lin_cols;  // device: slot of the column in E (one per coefficient-1 linear monomial)
	const uint32_t *comp_pts;  // device [n_comp]
	const uint8_t *lag;        // device [n_pts][128]
	const uint4 *E;            // device [n_slots][128]
	uint4 *out;
	uint32_t n_out;
};
__global__ void __launch_bounds__(128) k_uni_linear(const uint8_t *__restrict__ g_tables, const LinArgs A) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	__shared__ uint4 Ec[128];
	const uint32_t c = blockIdx.x, t = threadIdx.x, n_in = A.comp_pts[c];
	const uint32_t j0 = A.lin_off[c], j1 = A.lin_off[c + 1];
	if (j0 == j1 || n_in == 0) return;
	uint4 e = u4_zero();
	for (uint32_t j = j0; j < j1; j++) e ^= A.E[(uint64_t)A.lin_cols[j] * 128 + t];
	Ec[t] = e;
	__syncthreads();
	for (uint32_t i = t; i < n_in; i += 128) {
		const uint8_t *row = A.lag + (uint64_t)i * 128;
		uint4 acc = u4_zero();
		for (uint32_t u = 0; u < 128; u++) acc ^= f_mul128_b8_uniform(T, Ec[u], row[u]);  // (Ec[u] is uniform in the warp, the coefficient varies)
		A.out[(uint64_t)c * A.n_out + i] ^= acc;  // the only writer of this element in this launch; stream-ordered after k_uni_b8
	}
}

// extrapolate_round_evals (univariate.rs:565-640): composition c was evaluated at n_in = (deg_c-1)*2^k
// points; with zeros on the skipped domain those values determine a polynomial of degree < deg_c*2^k,
// whose values on the rest of the domain are B8-linear combinations of the evaluations:
//   out[c][i] = sum_{t < n_in} E_c[i - n_in][t] * out[c][t]        (i >= n_in)
struct ExtArgs {
	const uint32_t *comp_pts;  // device [n_comp]
	const uint32_t *ext_off;   // device [n_comp]: byte offset of E_c in `ext`
	const uint8_t *ext;
	uint4 *out;
	uint32_t n_out;
};
__global__ void __launch_bounds__(256) k_uni_extend(const uint8_t *__restrict__ g_tables, const ExtArgs A) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	const uint32_t c = blockIdx.x, n_in = A.comp_pts[c];
	if (n_in == 0 || n_in >= A.n_out) return;
	uint4 *row = A.out + (uint64_t)c * A.n_out;
	const uint8_t *E = A.ext + A.ext_off[c];
	for (uint32_t i = n_in + threadIdx.x; i < A.n_out; i += blockDim.x) {
		uint4 acc = u4_zero();
		for (uint32_t t = 0; t < n_in; t++) acc ^= f_mul128_b8_uniform(T, row[t], E[(uint64_t)(i - n_in) * n_in + t]);  // (row[t] is uniform in the warp)
		row[i] = acc;
	}
}

}  // namespace uni
}  // namespace b200
