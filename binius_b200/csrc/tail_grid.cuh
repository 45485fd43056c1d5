// tail_grid.cuh -- the persistent eq-ind sumcheck on a CO-RESIDENT GRID: every round of a zerocheck / eq-ind
// sumcheck whose data fits the caches (BASELINE config #3: 5 multilinears of 18 variables = 21 MB) in ONE kernel.
//
// Same contract as k_sumcheck_tail (kernels.cuh; reference per round = sumcheck_compute_round_evals
// hal/src/sumcheck_round_calculation.rs:126-349 + fold_left_lerp_inplace math/src/fold.rs:648-696 +
// fold_partial_eq_ind core/src/protocols/sumcheck/prove/common.rs:60-68), but the hypercube is spread over G CTAs
// (cooperative launch, one per SM):
//   * the hypercube indices are cut into chunks of 32 (one warp-load of 16-byte elements); chunk c belongs to CTA
//     c mod G IN EVERY ROUND.  HighToLow folding pairs i with i + half, so while half >= 32 G both elements of a pair,
//     the folded element and the eq-indicator entry have the same owner: a round then needs ONE grid barrier (before
//     CTA 0 posts the XOR-combined round values) and no data crosses between SMs.  Smaller rounds add a second barrier
//     after the fold, and CTAs that have run out of chunks leave (the barrier counts are a fixed function of the
//     round, so every CTA computes the same targets).
//   * inside a CTA the (value, chunk) items of a round are dealt to the warps; a warp reduces with shuffles into a
//     shared-memory accumulator (64-bit XOR atomics), the CTA adds its accumulators to the global ones.
//   * CTA 0 owns the host-mapped mailbox (values out, challenge in) and re-publishes the challenge in device memory.
// All multilinear / eq-indicator loads bypass L1 (ld.global.cg): other SMs wrote them in the previous round.
#pragma once
#include "kernels.cuh"

namespace b200 {

constexpr uint32_t TG_THREADS = 1024, TG_MAX_VALS = 1024, TG_MAX_CTAS = 128;

struct TailGridArgs {
	TailArgs t;
	uint4 *acc;                // device [n_vars][n_vals], zeroed before the launch
	uint32_t *bar;             // device: arrival counter of the grid barriers, zeroed
	uint4 *g_chal;             // device [n_vars]: challenge r as re-published by CTA 0
	uint32_t *g_chal_seq;      // device [n_vars], zeroed: == r + 1 when g_chal[r] is valid
	uint32_t *g_abort;         // device, zeroed: a watchdog expired somewhere
	uint32_t first_skip;       // the first round does not need the first `first_skip` evaluation points
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
	uint32_t v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// C^{(p)}(P(i)) for full-length multilinears, operands through L2
__device__ __forceinline__ uint4 tg_eval_point(const FieldTables &T, uint4 *const *mls, uint64_t half, const DevExpr &E, uint32_t code, uint4 z, uint64_t i) {
	uint4 tmp[MAX_EXPR_STEPS];
	for (uint32_t s = 0; s < E.n_steps; s++) {
		const b200_expr_step st = E.steps[s];
		uint4 v;
		switch (st.op) {
		case 0: v = tmp[st.l] ^ tmp[st.r]; break;
		case 1: v = f_mul128(T, tmp[st.l], tmp[st.r]); break;
		case 2: v = f_pow128(T, tmp[st.l], st.r); break;
		case 3: v = make_uint4((uint32_t)st.c_lo, (uint32_t)(st.c_lo >> 32), (uint32_t)st.c_hi, (uint32_t)(st.c_hi >> 32)); break;
		default: {
			const uint4 *m = mls[st.l];
			const uint4 hi = __ldcg(m + half + i);
			if (code == 1) v = hi;
			else {
				const uint4 lo = __ldcg(m + i);
				const uint4 d = hi ^ lo;
				v = code == 2 ? d : (lo ^ f_mul128(T, d, z));
			}
		}
		}
		tmp[s] = v;
	}
	return E.n_steps ? tmp[E.n_steps - 1] : u4_zero();
}

__global__ void __launch_bounds__(TG_THREADS, 1) k_sumcheck_tail_grid(const uint8_t *__restrict__ g_tables, const TailGridArgs GA) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	unsigned long long *acc_s = reinterpret_cast<unsigned long long *>(smem + ((FIELD_TABLE_BYTES + 127) & ~127u));  // [n_vals][2]
	__shared__ uint4 z_s;
	__shared__ uint32_t abort_s;
	const TailArgs &A = GA.t;
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
	const uint32_t G = gridDim.x, k = blockIdx.x;
	const uint32_t n_vals = A.n_comp * A.n_points;
	uint32_t bar_target = 0;
	if (tid == 0) abort_s = 0;
	__syncthreads();

	// grid barrier among the CTAs 0..P-1 (all of which are resident: cooperative launch)
	auto barrier = [&](uint32_t P) {
		bar_target += P;
		__syncthreads();
		if (tid == 0) {
			__threadfence();
			atomicAdd(GA.bar, 1u);
			const uint64_t t0 = globaltimer_ns();
			uint32_t spins = 0;
			while (ld_acquire_u32(GA.bar) < bar_target) {
				if ((++spins & 255u) == 0) {
					if (ld_acquire_u32(GA.g_abort)) {
						abort_s = 1;
						break;
					}
					if (globaltimer_ns() - t0 > A.timeout_ns) {
						atomicExch(GA.g_abort, 1u);
						abort_s = 1;
						break;
					}
				}
			}
			__threadfence();
		}
		__syncthreads();
	};

	for (uint32_t r = 0; r < A.n_vars; r++) {
		const uint64_t half = 1ull << (A.n_vars - 1 - r);
		const uint32_t n_chunks = (uint32_t)((half + 31) >> 5);
		const uint32_t P = n_chunks < G ? n_chunks : G;  // CTAs that take part in this round (k < P)
		const uint32_t my_chunks = (n_chunks - k + G - 1) / G;  // chunks k, k + G, ...
		// ---- round values
		for (uint32_t t = tid; t < 2 * n_vals; t += blockDim.x) acc_s[t] = 0;
		__syncthreads();
		for (uint32_t q = warp; q < n_vals * my_chunks; q += nw) {
			const uint32_t j = q / my_chunks, lc = q - j * my_chunks;
			const uint32_t c = j / A.n_points, p = j - c * A.n_points, code = A.codes[p];
			if (r == 0 && p < GA.first_skip) continue;
			const DevExpr X = code == 2 ? A.leads[c] : A.comps[c];
			if (X.n_steps == 0) continue;
			if (X.n_steps == 1) {
				const b200_expr_step st = X.steps[0];
				if (st.op == 3 && (st.c_lo | st.c_hi) == 0) continue;  // the constant 0 (a linear composition has no leading term)
			}
			const uint64_t i = 32ull * (k + (uint64_t)G * lc) + lane;
			uint4 v = u4_zero();
			if (i < half) {
				v = tg_eval_point(T, A.mls, half, X, code, A.points[p], i);
				v = f_mul128(T, v, __ldcg(A.eq_ind + i));
			}
			v = warp_xor(v);
			if (lane == 0) {
				const unsigned long long lo = v.x | ((unsigned long long)v.y << 32), hi = v.z | ((unsigned long long)v.w << 32);
				if (lo) atomicXor(acc_s + 2 * j, lo);
				if (hi) atomicXor(acc_s + 2 * j + 1, hi);
			}
		}
		__syncthreads();
		if (P > 1) {
			unsigned long long *acc_g = reinterpret_cast<unsigned long long *>(GA.acc + (uint64_t)r * n_vals);
			for (uint32_t t = tid; t < 2 * n_vals; t += blockDim.x)
				if (acc_s[t]) atomicXor(acc_g + t, acc_s[t]);
			barrier(P);
			if (abort_s) break;
		}
		// ---- values out, challenge in
		if (k == 0) {
			for (uint32_t t = tid; t < n_vals; t += blockDim.x) {
				uint4 v;
				if (P > 1) v = __ldcg(GA.acc + (uint64_t)r * n_vals + t);
				else v = make_uint4((uint32_t)acc_s[2 * t], (uint32_t)(acc_s[2 * t] >> 32), (uint32_t)acc_s[2 * t + 1], (uint32_t)(acc_s[2 * t + 1] >> 32));
				volatile uint4 *dst = A.mb_vals + (uint64_t)r * n_vals + t;
				dst->x = v.x, dst->y = v.y, dst->z = v.z, dst->w = v.w;
				__threadfence_system();
			}
			__syncthreads();
			if (tid == 0) {
				A.mb_seq[r] = r + 1;
				__threadfence_system();
				const uint64_t t0 = globaltimer_ns();
				while (A.mb_chal_seq[r] != r + 1) {
					if (globaltimer_ns() - t0 > A.timeout_ns) {
						abort_s = 1;
						break;
					}
				}
				if (abort_s) atomicExch(GA.g_abort, 1u);
				else {
					volatile uint4 *zc = A.mb_chal + r;
					const uint4 z = make_uint4(zc->x, zc->y, zc->z, zc->w);
					z_s = z;
					if (P > 1) {
						GA.g_chal[r] = z;
						__threadfence();
						st_release_u32(GA.g_chal_seq + r, r + 1);
					}
				}
			}
		} else if (tid == 0) {
			const uint64_t t0 = globaltimer_ns();
			uint32_t spins = 0;
			while (ld_acquire_u32(GA.g_chal_seq + r) != r + 1) {
				__nanosleep(40);
				if ((++spins & 255u) == 0) {
					if (ld_acquire_u32(GA.g_abort)) {
						abort_s = 1;
						break;
					}
					if (globaltimer_ns() - t0 > 2 * A.timeout_ns) {
						atomicExch(GA.g_abort, 1u);
						abort_s = 1;
						break;
					}
				}
			}
			if (!abort_s) z_s = __ldcg(GA.g_chal + r);
		}
		__syncthreads();
		if (abort_s) break;
		const uint4 z = z_s;
		// ---- fold the multilinears, halve the eq-indicator (own chunks only)
		for (uint32_t q = warp; q < A.m * my_chunks; q += nw) {
			const uint32_t t = q / my_chunks, lc = q - t * my_chunks;
			const uint64_t i = 32ull * (k + (uint64_t)G * lc) + lane;
			if (i < half) {
				uint4 *ml = A.mls[t];
				const uint4 lo = __ldcg(ml + i), hi = __ldcg(ml + half + i);
				ml[i] = lo ^ f_mul128(T, lo ^ hi, z);
			}
		}
		const uint64_t hn = half >> 1;
		for (uint32_t lc = warp; lc < my_chunks; lc += nw) {
			const uint64_t i = 32ull * (k + (uint64_t)G * lc) + lane;
			if (i < hn) A.eq_ind[i] = __ldcg(A.eq_ind + i) ^ __ldcg(A.eq_ind + hn + i);
		}
		if (r + 1 == A.n_vars) break;
		// next round pairs i with i + hn: same owner iff hn is a multiple of 32 G
		const bool local = hn >= 32ull * G;
		if (P > 1 && !local) {
			barrier(P);
			if (abort_s) break;
		} else __syncthreads();
		const uint32_t n_next = (uint32_t)((hn + 31) >> 5);
		if (k >= (n_next < G ? n_next : G)) return;
	}
	if (abort_s && tid == 0) {
		*A.status = 1;
		__threadfence_system();
	}
}

}  // namespace b200
