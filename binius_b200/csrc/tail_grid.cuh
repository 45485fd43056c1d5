// tail_grid.cuh -- the persistent eq-ind sumcheck on a CO-RESIDENT GRID: every round of a zerocheck / eq-ind
// sumcheck whose data fits the caches (BASELINE config #3: 5 multilinears of 18 variables = 21 MB) in ONE kernel.
//
// Late sumcheck rounds are a few kilobytes of data and through separate calls they are bounded by launch + copy +
// synchronisation latency (~40 us per round); here ONE kernel loops over the rounds: round values (semantics of
// k_eq_ind_round_evals, HighToLow, full-length multilinears; reference sumcheck_compute_round_evals
// hal/src/sumcheck_round_calculation.rs:126-349) -> host-mapped mailbox -> challenge from the host-mapped mailbox ->
// fold_left_lerp_inplace (math/src/fold.rs:648-696) of every multilinear + fold_partial_eq_ind (core/src/protocols/
// sumcheck/prove/common.rs:60-68).  A watchdog on %globaltimer ends the kernel (status = 1) if no challenge arrives.
// The hypercube is spread over G CTAs (cooperative launch, one per SM; G = 1 for instances of <= 32 points):
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

constexpr uint32_t TG_THREADS = 1024, TG_MAX_VALS = 4096, TG_MAX_CTAS = 128, TG_STEP_CACHE_BYTES = 64u << 10, TG_SMEM_MAX = 224u << 10;  // dynamic; the kernel has 2 KiB of static shared memory

// Mailbox protocol (host-mapped memory, no fences on the critical path): every 16-byte datum d travels as the 32-byte
// pair {d, ~d} into a slot that the host zeroed before the launch.  A reader polls the slot until the two halves are
// complements: 8-byte pieces arrive in any order, each piece is either still zero or final, and a pair of pieces
// (x, y) with x == ~y can only be (final, final) -- or x happens to equal its final value already.  One PCIe read
// (two lanes x 16 B) therefore returns a validated challenge, and the device never waits for a posted write.
struct TailArgs {
	uint4 *const *mls;   // device [m]
	uint32_t m, n_vars;  // n_vars rounds remain
	uint4 *eq_ind;       // 2^(n_vars - 1), halved in place
	const DevExpr *comps, *leads;
	uint32_t n_comp, n_points;
	const uint32_t *codes;
	const uint4 *points;
	uint4 *mb_vals;            // host-mapped [n_vars][n_comp * n_points][2]
	uint4 *mb_chal;            // host-mapped [n_vars][2]
	volatile uint32_t *status; // host-mapped: 0 running / done, 1 watchdog expired
	uint64_t *mb_trace;        // host-mapped [n_vars][4] %globaltimer stamps of CTA 0, or null
	uint64_t timeout_ns;
};
struct TailGridArgs {
	TailArgs t;
	uint4 *acc;                // device [n_vars][n_vals], zeroed before the launch
	uint32_t *bar;             // device: arrival counter of the grid barriers, zeroed
	uint4 *g_chal;             // device [n_vars]: challenge r as re-published by CTA 0
	uint32_t *g_chal_seq;      // device [n_vars], zeroed: == r + 1 when g_chal[r] is valid
	uint32_t *g_abort;         // device, zeroed: a watchdog expired somewhere
	uint32_t first_skip;       // the first round does not need the first `first_skip` evaluation points
	const uint32_t *step_off;  // device [2 n_comp + 1] prefix sums of the step counts (comps, then leads), or null: no cache
	uint32_t off_acc, off_ex, off_steps;  // shared-memory layout (bytes)
	uint32_t off_hdr;   // multilinear pointers [m], point codes [n_points], points [n_points] (0: read from global memory)
	uint32_t off_k64;   // 64 KiB + 1.5 KiB for the table of x -> x * challenge (linmap.cuh K64 engine), 0: none
	uint32_t res_half;  // from the round with this many (or fewer) index pairs on, CTA 0 keeps the multilinears and the
						// eq-indicator in shared memory (in the K64 region, which only the large rounds use) and runs
						// alone: no barriers, no L2 latency in the dependent chain of the small rounds.  0: never
};

__device__ __forceinline__ uint64_t globaltimer_ns() {
	uint64_t t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
	uint32_t v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_volatile_u4(uint4 *p, uint4 v) {
	asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_u4(const uint4 *p) {
	uint4 v;
	asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
	return v;
}

// C^{(p)}(P(i)) for full-length multilinears, operands through L2 (RES: from the shared-memory copies)
template <bool RES>
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
			const uint4 hi = RES ? m[half + i] : __ldcg(m + half + i);
			if (code == 1) v = hi;
			else {
				const uint4 lo = RES ? m[i] : __ldcg(m + i);
				const uint4 d = hi ^ lo;
				v = code == 2 ? d : (lo ^ f_mul128(T, z, d));
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
	unsigned long long *acc_s = reinterpret_cast<unsigned long long *>(smem + GA.off_acc);  // [n_vals][2]
	__shared__ uint4 z_s;
	__shared__ uint32_t abort_s;
	const TailArgs &A = GA.t;
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
	const uint32_t k = blockIdx.x;
	const uint32_t n_vals = A.n_comp * A.n_points;
	uint32_t bar_target = 0;
	if (tid == 0) abort_s = 0;
	// the expressions (comps, then leads), with their steps in shared memory when they fit: the interpreter reads a
	// step per operation, and in the small rounds that latency is the round
	const DevExpr *comps = A.comps, *leads = A.leads;
	uint4 *const *mls = A.mls;
	const uint32_t *codes = A.codes;
	const uint4 *points = A.points;
	if (GA.off_hdr) {
		uint4 *pt_s = reinterpret_cast<uint4 *>(smem + GA.off_hdr);
		uint4 **ml_s = reinterpret_cast<uint4 **>(pt_s + A.n_points);
		uint32_t *cd_s = reinterpret_cast<uint32_t *>(ml_s + A.m);
		for (uint32_t t = tid; t < A.n_points; t += blockDim.x) pt_s[t] = A.points[t], cd_s[t] = A.codes[t];
		for (uint32_t t = tid; t < A.m; t += blockDim.x) ml_s[t] = A.mls[t];
		mls = ml_s, codes = cd_s, points = pt_s;
	}
	if (GA.step_off) {
		DevExpr *ex_s = reinterpret_cast<DevExpr *>(smem + GA.off_ex);
		b200_expr_step *st_s = reinterpret_cast<b200_expr_step *>(smem + GA.off_steps);
		for (uint32_t e = warp; e < 2 * A.n_comp; e += nw) {
			const DevExpr src = e < A.n_comp ? A.comps[e] : A.leads[e - A.n_comp];
			const uint32_t off = GA.step_off[e];
			for (uint32_t s = lane; s < 2 * src.n_steps; s += 32) reinterpret_cast<uint4 *>(st_s + off)[s] = reinterpret_cast<const uint4 *>(src.steps)[s];
			if (lane == 0) {
				DevExpr d = src;
				d.steps = st_s + off;
				ex_s[e] = d;
			}
		}
		comps = ex_s, leads = ex_s + A.n_comp;
	}
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
	const bool trace = A.mb_trace != nullptr && k == 0 && tid == 0;

	bool res = false;      // CTA 0 runs alone on its shared-memory copies
	uint4 *eq = A.eq_ind;
	uint32_t G = gridDim.x;
	for (uint32_t r = 0; r < A.n_vars; r++) {
		const uint64_t half = 1ull << (A.n_vars - 1 - r);
		if (!res && GA.res_half && half <= GA.res_half) {
			// every fold of the previous round is visible (that round ended with a grid barrier: half < 32 G)
			if (k != 0) return;
			uint4 *base = reinterpret_cast<uint4 *>(smem + GA.off_k64);
			uint4 **ml_s = const_cast<uint4 **>(mls);  // the pointer table in the shared-memory header (GA.off_hdr != 0 when res_half != 0)
			for (uint32_t t = warp; t < A.m; t += nw) {
				const uint4 *src = mls[t];
				for (uint32_t i = lane; i < 2 * half; i += 32) base[(uint64_t)t * 2 * half + i] = __ldcg(src + i);
			}
			for (uint32_t i = tid; i < half; i += blockDim.x) base[(uint64_t)A.m * 2 * half + i] = __ldcg(A.eq_ind + i);
			__syncthreads();
			for (uint32_t t = tid; t < A.m; t += blockDim.x) ml_s[t] = base + (uint64_t)t * 2 * half;
			eq = base + (uint64_t)A.m * 2 * half;
			res = true, G = 1;
			__syncthreads();
		}
		const uint32_t n_chunks = (uint32_t)((half + 31) >> 5);
		const uint32_t P = n_chunks < G ? n_chunks : G;  // CTAs that take part in this round (k < P)
		const uint32_t my_chunks = (n_chunks - k + G - 1) / G;  // chunks k, k + G, ...
		if (trace) A.mb_trace[4 * r] = globaltimer_ns();
		// ---- round values
		for (uint32_t t = tid; t < 2 * n_vals; t += blockDim.x) acc_s[t] = 0;
		__syncthreads();
		for (uint32_t q = warp; q < n_vals * my_chunks; q += nw) {
			const uint32_t j = q / my_chunks, lc = q - j * my_chunks;
			const uint32_t c = j / A.n_points, p = j - c * A.n_points, code = codes[p];
			if (r == 0 && p < GA.first_skip) continue;
			const DevExpr X = code == 2 ? leads[c] : comps[c];
			if (X.n_steps == 0) continue;
			if (X.n_steps == 1) {
				const b200_expr_step st = X.steps[0];
				if (st.op == 3 && (st.c_lo | st.c_hi) == 0) continue;  // the constant 0 (a linear composition has no leading term)
			}
			const uint64_t i = 32ull * (k + (uint64_t)G * lc) + lane;
			uint4 v = u4_zero();
			if (i < half) {
				v = res ? tg_eval_point<true>(T, mls, half, X, code, points[p], i) : tg_eval_point<false>(T, mls, half, X, code, points[p], i);
				v = f_mul128(T, v, res ? eq[i] : __ldcg(eq + i));
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
		if (trace) A.mb_trace[4 * r + 1] = globaltimer_ns();
		// ---- values out, challenge in
		if (k == 0) {
			for (uint32_t t = tid; t < n_vals; t += blockDim.x) {
				uint4 v;
				if (P > 1) v = __ldcg(GA.acc + (uint64_t)r * n_vals + t);
				else v = make_uint4((uint32_t)acc_s[2 * t], (uint32_t)(acc_s[2 * t] >> 32), (uint32_t)acc_s[2 * t + 1], (uint32_t)(acc_s[2 * t + 1] >> 32));
				uint4 *dst = A.mb_vals + 2 * ((uint64_t)r * n_vals + t);
				st_volatile_u4(dst, v);
				st_volatile_u4(dst + 1, make_uint4(~v.x, ~v.y, ~v.z, ~v.w));
			}
			if (warp == 0) {
				const uint64_t t0 = globaltimer_ns();
				const uint4 *src = A.mb_chal + 2 * r + (lane & 1);
				uint4 z;
				bool ok = false;
				for (uint32_t spins = 0; !ok; spins++) {
					const uint4 w = lane < 2 ? ld_volatile_u4(src) : u4_zero();
					z = make_uint4(__shfl_sync(0xffffffffu, w.x, 0), __shfl_sync(0xffffffffu, w.y, 0), __shfl_sync(0xffffffffu, w.z, 0), __shfl_sync(0xffffffffu, w.w, 0));
					const uint4 zc = make_uint4(__shfl_sync(0xffffffffu, w.x, 1), __shfl_sync(0xffffffffu, w.y, 1), __shfl_sync(0xffffffffu, w.z, 1), __shfl_sync(0xffffffffu, w.w, 1));
					ok = z.x == ~zc.x && z.y == ~zc.y && z.z == ~zc.z && z.w == ~zc.w;
					if (!ok && (spins & 15u) == 15u && globaltimer_ns() - t0 > A.timeout_ns) break;
				}
				if (lane == 0) {
					if (!ok) {
						abort_s = 1;
						atomicExch(GA.g_abort, 1u);
					} else {
						z_s = z;
						if (P > 1) {
							GA.g_chal[r] = z;
							__threadfence();
							st_release_u32(GA.g_chal_seq + r, r + 1);
						}
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
		if (trace) A.mb_trace[4 * r + 2] = globaltimer_ns();
		const uint4 z = z_s;
		// ---- fold the multilinears, halve the eq-indicator (own chunks only)
		// (two or more items per warp pay for the ~2 us build of the challenge's K64 table: 24 conflict-free LDS.64 per
		// product instead of the 121 byte lookups of the general multiply)
		if (GA.off_k64 && !res && A.m * my_chunks >= 2 * nw) {
			k64_build_mul(smem + GA.off_k64, reinterpret_cast<uint2 *>(smem + GA.off_k64 + LUT_BYTES), z);
			const K64Lane L = k64_lane_init(smem + GA.off_k64);
			for (uint32_t q = warp; q < A.m * my_chunks; q += nw) {
				const uint32_t t = q / my_chunks, lc = q - t * my_chunks;
				const uint64_t i = 32ull * (k + (uint64_t)G * lc) + lane;
				if (i < half) {
					uint4 *ml = mls[t];
					const uint4 lo = __ldcg(ml + i), hi = __ldcg(ml + half + i);
					ml[i] = lo ^ k64_apply(L, lo ^ hi);
				}
			}
		} else {
			for (uint32_t q = warp; q < A.m * my_chunks; q += nw) {
				const uint32_t t = q / my_chunks, lc = q - t * my_chunks;
				const uint64_t i = 32ull * (k + (uint64_t)G * lc) + lane;
				if (i < half) {
					uint4 *ml = mls[t];
					const uint4 lo = res ? ml[i] : __ldcg(ml + i), hi = res ? ml[half + i] : __ldcg(ml + half + i);
					ml[i] = lo ^ f_mul128(T, z, lo ^ hi);  // warp-uniform operand first (field.cuh)
				}
			}
		}
		const uint64_t hn = half >> 1;
		for (uint32_t lc = warp; lc < my_chunks; lc += nw) {
			const uint64_t i = 32ull * (k + (uint64_t)G * lc) + lane;
			if (i < hn) eq[i] = res ? eq[i] ^ eq[hn + i] : __ldcg(eq + i) ^ __ldcg(eq + hn + i);
		}
		if (r + 1 == A.n_vars) {
			if (res) {  // the final folds go back to device memory (nothing in between is observable: the context refuses calls)
				__syncthreads();
				for (uint32_t t = tid; t < A.m; t += blockDim.x) A.mls[t][0] = mls[t][0];
				if (tid == 0) A.eq_ind[0] = eq[0];
			}
			break;
		}
		// next round pairs i with i + hn: same owner iff hn is a multiple of 32 G
		const bool local = hn >= 32ull * G;
		if (P > 1 && !local) {
			barrier(P);
			if (abort_s) break;
		} else __syncthreads();
		if (trace) A.mb_trace[4 * r + 3] = globaltimer_ns();
		const uint32_t n_next = (uint32_t)((hn + 31) >> 5);
		if (k >= (n_next < G ? n_next : G)) return;
	}
	if (abort_s && tid == 0) {
		*A.status = 1;
		__threadfence_system();
	}
}

}  // namespace b200
