// kernels.cuh -- CUDA kernels (sm_100a) for the ComputeLayer / ComputationBackend hot path.
// Host-side launch wrappers + validation live in capi.cu.  Reference citations are next to each
// kernel; semantics are restated for the checker in oracle/ops.c.
#pragma once
#include "field.cuh"
#include "linmap.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------------
// fold-high / extrapolate_line over a list of segments sharing ONE challenge z:
//     e0[i] ^= (e1[i] ^ e0[i]) * z            i <  pivot
//     e0[i] ^= (suffix ^ e0[i]) * z           pivot <= i < upper
// reference: compute/src/cpu/layer.rs:393-408 (extrapolate_line) and math/src/fold.rs:648-696
// (fold_left_lerp_inplace with const-suffix; packed width 1).
struct LerpSeg {
	uint4 *e0;
	const uint4 *e1;
	uint64_t pivot;
	uint64_t upper;
	uint4 suffix;
	uint64_t tile_start;  // first global tile of this segment (prefix sum), filled by the host
};
constexpr uint32_t LERP_MAX_SEGS = 48;  // segments passed by value in the kernel parameters
struct LerpArgs {
	LerpSeg segs[LERP_MAX_SEGS];
	const LerpSeg *segs_dev;  // more than LERP_MAX_SEGS segments: the list staged in device memory (else null)
	uint32_t n_segs;
	uint64_t n_tiles;
	uint4 z;
};

constexpr uint32_t FOLD_THREADS = 512;
constexpr uint32_t FOLD_UNROLL = 2;
constexpr uint32_t FOLD_TILE = FOLD_THREADS * FOLD_UNROLL;

// The two constant-multiplication engines of linmap.cuh behind one interface
template <bool K64> struct MulEngine;
template <> struct MulEngine<false> {
	LutLane L;
	const uint8_t *tbl;
	__device__ __forceinline__ MulEngine(uint8_t *smem, uint4 z) : tbl(smem) {
		lut_build_mul(smem, reinterpret_cast<uint4 *>(smem + LUT_BYTES), z);
		L = lut_lane_init();
	}
	__device__ __forceinline__ uint4 mul(uint4 x) const { return lut_apply(tbl, L, x); }
};
template <> struct MulEngine<true> {
	K64Lane L;
	const uint8_t *tbl;
	__device__ __forceinline__ MulEngine(uint8_t *smem, uint4 z) : tbl(smem) {
		k64_build_mul(smem, reinterpret_cast<uint2 *>(smem + LUT_BYTES), z);
		L = k64_lane_init(smem);
	}
	__device__ __forceinline__ uint4 mul(uint4 x) const { return k64_apply(L, x); }
};

// THREADS x UNR elements per tile; host must compute tile_start with the same tile size
template <uint32_t THREADS, uint32_t UNR, int MINB, bool K64>
__global__ void __launch_bounds__(THREADS, MINB) k_lerp_lut(const __grid_constant__ LerpArgs A) {
	extern __shared__ __align__(256) uint8_t smem[];
	const MulEngine<K64> E(smem, A.z);
	constexpr uint32_t TILE = THREADS * UNR;

	uint32_t seg = 0;
	for (uint64_t tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
		while (seg + 1 < A.n_segs && A.segs[seg + 1].tile_start <= tile) seg++;
		const LerpSeg &S = A.segs[seg];
		uint64_t base = (tile - S.tile_start) * TILE + threadIdx.x;
		const uint64_t upper = S.upper, pivot = S.pivot;
		uint4 *e0 = S.e0;
		const uint4 *e1 = S.e1;
		uint4 a[UNR], b[UNR];
#pragma unroll
		for (uint32_t u = 0; u < UNR; u++) {
			uint64_t i = base + (uint64_t)u * THREADS;
			if (i < upper) {
				a[u] = e0[i];
				b[u] = i < pivot ? __ldg(e1 + i) : S.suffix;
			}
		}
#pragma unroll
		for (uint32_t u = 0; u < UNR; u++) {
			uint64_t i = base + (uint64_t)u * THREADS;
			if (i < upper) e0[i] = a[u] ^ E.mul(a[u] ^ b[u]);
		}
	}
}

// LowToHigh fold (fold_right_lerp, math/src/fold.rs:528-575, as driven by hal/src/sumcheck_folding.rs:
// 117-143): out[i] = in[2i] + (in[2i+1] - in[2i]) * z for i < pivot (= prefix / 2); the odd tail
// element pairs with the constant suffix.  Same LerpArgs: e0 = out, e1 = in, upper = ceil(prefix / 2).
template <uint32_t THREADS, uint32_t UNR, int MINB, bool K64>
__global__ void __launch_bounds__(THREADS, MINB) k_lerp_pairs_lut(const __grid_constant__ LerpArgs A) {
	extern __shared__ __align__(256) uint8_t smem[];
	const MulEngine<K64> E(smem, A.z);
	constexpr uint32_t TILE = THREADS * UNR;

	uint32_t seg = 0;
	for (uint64_t tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
		while (seg + 1 < A.n_segs && A.segs[seg + 1].tile_start <= tile) seg++;
		const LerpSeg &S = A.segs[seg];
		uint64_t base = (tile - S.tile_start) * TILE + threadIdx.x;
		const uint64_t upper = S.upper, pivot = S.pivot;
		uint4 *out = S.e0;
		const uint4 *in = S.e1;
		uint4 a[UNR], b[UNR];
#pragma unroll
		for (uint32_t u = 0; u < UNR; u++) {
			uint64_t i = base + (uint64_t)u * THREADS;
			if (i < upper) {
				a[u] = __ldg(in + 2 * i);
				b[u] = i < pivot ? __ldg(in + 2 * i + 1) : S.suffix;
			}
		}
#pragma unroll
		for (uint32_t u = 0; u < UNR; u++) {
			uint64_t i = base + (uint64_t)u * THREADS;
			if (i < upper) out[i] = a[u] ^ E.mul(a[u] ^ b[u]);
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Tensor expansion, R <= 3 rounds fused per launch:  round t:  p = x * r_t ; lo = x ^ p ; hi = p.
// Each thread takes one input x[i] (i < n0), keeps its 2^R descendants in registers and stores them at
// i + e * n0 (in place: a thread only touches its own column), so the intermediate rounds never travel
// to HBM.  One Karatsuba-64 table per coordinate (linmap.cuh), R x 64 KiB of shared memory.
// reference: compute/src/layer.rs:269-296 (definition), math/src/tensor_prod_eq_ind.rs:35-77
constexpr uint32_t EXP_THREADS = 512;
struct ExpandArgs {
	uint4 *data;
	uint64_t n0;
	uint4 r[3];
};
template <uint32_t R>
__global__ void __launch_bounds__(EXP_THREADS, 1) k_expand_k64(const __grid_constant__ ExpandArgs A) {
	extern __shared__ __align__(256) uint8_t smem[];
	uint2 *stage = reinterpret_cast<uint2 *>(smem + R * LUT_BYTES);
	uint4 zs[R];
#pragma unroll
	for (uint32_t t = 0; t < R; t++) zs[t] = A.r[t];
	k64_build_multi<R>(smem, stage, zs);
	K64Lane L = k64_lane_init(smem);
	const uint32_t sbase = L.sbase;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < A.n0; i += (uint64_t)gridDim.x * blockDim.x) {
		uint4 y[1u << R];
		y[0] = A.data[i];
#pragma unroll
		for (uint32_t t = 0; t < R; t++) {
			L.sbase = sbase + t * LUT_BYTES;
#pragma unroll
			for (uint32_t e = 0; e < (1u << t); e++) {
				const uint4 p = k64_apply(L, y[e]);
				y[e] ^= p;
				y[e + (1u << t)] = p;
			}
		}
#pragma unroll
		for (uint32_t e = 0; e < (1u << R); e++) A.data[i + e * A.n0] = y[e];
	}
}

// small tensor expansion: rounds [0, k) entirely inside one CTA's shared memory (2^(log_n+k) <= 4096).
// One nibble-LUT (8 KiB) per coordinate, all built up front in parallel; then k rounds of in-smem
// doubling.  dyn smem = k * (NLUT_BYTES + 2048) + 16 * 2^(log_n+k)
constexpr uint32_t EXP_SMALL_LOG = 12;
__global__ void __launch_bounds__(1024) k_expand_small(uint4 *__restrict__ data, uint32_t log_n, const uint4 *__restrict__ coords, uint32_t k) {
	extern __shared__ __align__(128) uint8_t smem[];
	uint4 *buf = reinterpret_cast<uint4 *>(smem + k * NLUT_BYTES);
	uint4 *img = buf + (1u << (log_n + k));  // [k][128] basis images beta_i * r_t, 2 KiB per coordinate
	for (uint32_t e = threadIdx.x; e < k * 128; e += blockDim.x) img[e] = basis_image(coords[e >> 7], e & 127);
	__syncthreads();
	for (uint32_t e = threadIdx.x; e < k * 512; e += blockDim.x) nlut_fill_from_images(smem + (e >> 9) * NLUT_BYTES, img + (e >> 9) * 128, e & 511);
	const uint32_t n0 = 1u << log_n;
	for (uint32_t i = threadIdx.x; i < n0; i += blockDim.x) buf[i] = data[i];
	__syncthreads();
	const NLutLane L = nlut_lane_init();
	for (uint32_t r = 0; r < k; r++) {
		const uint32_t half = 1u << (log_n + r);
		const uint8_t *lut = smem + r * NLUT_BYTES;
		for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
			const uint4 x = buf[i];
			const uint4 p = nlut_apply(lut, L, x);
			buf[i] = x ^ p;
			buf[half + i] = p;
		}
		__syncthreads();
	}
	const uint32_t n = 1u << (log_n + k);
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) data[i] = buf[i];
}

// Tensor expansion as an OUTER PRODUCT.  After k rounds  out[j * N + i] = in[i] * T[j]  with T = the expansion of [1]
// by the same coordinates (the rounds only ever multiply, and multiplication is commutative), so the doubling chain --
// k dependent launches whose small links are pure latency -- collapses to two independent short chains and one wide
// kernel:
//   k_expand_pair : up to three SMALL expansions, one CTA per problem: in' = the data brought to 2^6 elements and the
//                   tables T_mid (<= 2^6 entries) and T_hi (<= 2^10) of the following coordinates, by general per-lane
//                   products.  Small on purpose: one SM does ~10^8 general products per second, 4096 of them are 35 us.
//   k_expand_outer: out[j * N + i] = in[i] * T[j]: a CTA takes table entries j, builds the K64 table of T[j] (~3 us)
//                   and streams the N <= 4096 staged elements through it.  Used twice: in'' = in' x T_mid (2^12
//                   elements), then out = in'' x T_hi.
// reference: compute/src/layer.rs:269-296 (definition), math/src/tensor_prod_eq_ind.rs:35-77
struct ExpandPairArgs {
	const uint4 *src[3];   // 2^log_n[p] input elements (null: the single element 1)
	uint4 *dst[3];         // 2^(log_n[p] + k[p]) outputs
	uint32_t log_n[3], k[3];
	uint4 coords[3][EXP_SMALL_LOG];
};
// T = the expansion of [1] by k coordinates is the tensor product of the k two-element tables [1 + c_r, c_r]; adjacent
// tables are merged pairwise (new[hi * |lo| + lo] = lo_tab[lo] * hi_tab[hi]), all merges of a level in parallel:
// ceil(log2 k) dependent product rounds instead of k (a per-lane product of one warp is ~2 us of latency, and with
// <= 4096 elements latency is all there is).  A last round multiplies by the input elements when there are any.
// grid = number of problems (1 or 2), block = 1024, dyn smem = field tables + 2 x 16 * 2^(log_n + k) (ping-pong)
__global__ void __launch_bounds__(1024) k_expand_pair(const uint8_t *__restrict__ g_tables, const __grid_constant__ ExpandPairArgs A) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	const uint32_t p = blockIdx.x, log_n = A.log_n[p], k = A.k[p], n0 = 1u << log_n;
	const uint32_t cap = 1u << (log_n + k);
	uint4 *cur = reinterpret_cast<uint4 *>(smem + ((FIELD_TABLE_BYTES + 127) & ~127u)), *nxt = cur + (cap > 2 * k ? cap : 2 * k);
	// level 0: k tables of one coordinate each, table t at [2t, 2t + 2)
	for (uint32_t e = threadIdx.x; e < 2 * k; e += blockDim.x) {
		const uint4 c = A.coords[p][e >> 1];
		cur[e] = (e & 1) ? c : make_uint4(c.x ^ 1u, c.y, c.z, c.w);
	}
	__syncthreads();
	// every table but the last covers `w` coordinates; the last covers `w_last` (1 <= w_last <= w)
	uint32_t n_tab = k, w = 1, w_last = 1;
	while (n_tab > 1) {
		const uint32_t n_new = (n_tab + 1) / 2, S = 1u << w, S2 = S * S;
		const bool odd = n_tab & 1;  // the last table has no partner: copied
		const uint32_t last_bits = odd ? w_last : w + w_last, last_off_new = (n_new - 1) * S2;
		const uint32_t total = last_off_new + (1u << last_bits);
		for (uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
			const uint32_t t = e < last_off_new ? e / S2 : n_new - 1, idx = e - t * S2;
			const uint32_t lo_off = 2 * t * S;
			if (t == n_new - 1 && odd) nxt[e] = cur[lo_off + idx];
			else nxt[e] = f_mul128(T, cur[lo_off + S + (idx >> w)], cur[lo_off + (idx & (S - 1))]);  // the slowly varying operand FIRST (field.cuh)
		}
		__syncthreads();
		uint4 *tmp = cur;
		cur = nxt, nxt = tmp;
		w_last = last_bits, w *= 2, n_tab = n_new;
	}
	// cur = T (2^k entries; the single entry 1 when k = 0)
	uint4 *dst = A.dst[p];
	if (A.src[p]) {
		// the input may be the head of the output (in-place expansion): stage it before anything is written
		for (uint32_t e = threadIdx.x; e < n0; e += blockDim.x) nxt[e] = A.src[p][e];
		__syncthreads();
		// first operand = the one that takes fewer distinct values inside a warp (field.cuh): the input element while
		// n0 <= 4 (a single input element is the same in EVERY lane), the table entry from n0 = 8 on
		const bool x_first = n0 <= 4;
		for (uint32_t e = threadIdx.x; e < cap; e += blockDim.x) {
			const uint4 x = nxt[e & (n0 - 1)];
			dst[e] = !k ? x : x_first ? f_mul128(T, x, cur[e >> log_n]) : f_mul128(T, cur[e >> log_n], x);
		}
	} else {
		for (uint32_t e = threadIdx.x; e < cap; e += blockDim.x) dst[e] = k ? cur[e] : u4_one();
	}
}
// grid <= n_j, block = 1024, dyn smem = LUT_BYTES + 2048 + 16 * n
__global__ void __launch_bounds__(1024, 1) k_expand_outer(const uint4 *__restrict__ in, uint32_t n, const uint4 *__restrict__ tab, uint32_t n_j,
																  uint4 *__restrict__ out) {
	extern __shared__ __align__(256) uint8_t smem[];
	uint2 *stage = reinterpret_cast<uint2 *>(smem + LUT_BYTES);
	uint4 *in_s = reinterpret_cast<uint4 *>(smem + LUT_BYTES + 2048);
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) in_s[i] = __ldg(in + i);
	for (uint32_t j = blockIdx.x; j < n_j; j += gridDim.x) {
		k64_build_mul(smem, stage, __ldg(tab + j));  // (its first barrier also orders the staging loop / the previous j's reads)
		const K64Lane L = k64_lane_init(smem);
		uint4 *o = out + (uint64_t)j * n;
		for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) o[i] = k64_apply(L, in_s[i]);
		__syncthreads();
	}
}
// (A variant with two CTAs of 512 threads per SM and the input read through L1 -- so that one CTA's table build overlaps
// the other's products -- measured the same 48 us for 1024 tables x 4096 elements: the stage is bound by issue slots,
// ~6.8 us per table and SM for the build plus 4096 K64 products, not by the latency of the build.)

// ------------------------------------------------------------------------------------------------
__global__ void k_fill(uint4 *__restrict__ dst, uint64_t n, uint4 v) {
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = v;
}
// KernelExecutor::add / add_assign (compute/src/cpu/layer.rs:516-548)
__global__ void k_add(const uint4 *__restrict__ a, const uint4 *__restrict__ b, uint4 *__restrict__ dst, uint64_t n) {
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = a[i] ^ b[i];
}
__global__ void k_set_slot(uint4 *slot, uint4 v) { *slot = v; }
__global__ void k_xor_slot(uint4 *slot, uint4 v) { *slot ^= v; }

// ------------------------------------------------------------------------------------------------
// inner_product(SubfieldSlice{a, lvl}, b) = sum_i sum_j b[i*L + j] * limb_j(a[i])
// reference: compute/src/cpu/layer.rs:205-236
__global__ void __launch_bounds__(256) k_inner_product(const uint8_t *__restrict__ g_tables, const uint4 *__restrict__ a,
													   uint32_t lvl, const uint4 *__restrict__ b, uint64_t n_b, uint4 *__restrict__ slot) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	__shared__ uint4 red[32];
	uint32_t logL = 7 - lvl;
	uint4 acc = u4_zero();
	for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < n_b; e += (uint64_t)gridDim.x * blockDim.x) {
		uint4 av = __ldg(a + (e >> logL));
		uint4 s = f_limb(av, lvl, (uint32_t)(e & ((1u << logL) - 1)));
		acc ^= f_mul128_sub(T, __ldg(b + e), s, lvl);
	}
	acc = block_xor(acc, red);
	if (threadIdx.x == 0) atomic_xor_u4(slot, acc);
}

// fold_left : out[i] = sum_j vec[j] * evals[j*rows + i]      (cpu/layer.rs:574-621)
// fold_right: out[i] = sum_j vec[j] * evals[i*cols + j]      (cpu/layer.rs:628-675)
// evals = limbs of `mat` at tower level lvl, flattened low limb first.
template <bool RIGHT>
__global__ void __launch_bounds__(256) k_fold_mat(const uint8_t *__restrict__ g_tables, const uint4 *__restrict__ mat, uint32_t lvl,
												  const uint4 *__restrict__ vec, uint32_t log_q, uint4 *__restrict__ out, uint64_t n_out) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	uint32_t logL = 7 - lvl;
	uint64_t nq = 1ull << log_q;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_out; i += (uint64_t)gridDim.x * blockDim.x) {
		uint4 acc = u4_zero();
		for (uint64_t j = 0; j < nq; j++) {
			uint64_t e = RIGHT ? (i * nq + j) : (j * n_out + i);
			uint4 m = __ldg(mat + (e >> logL));
			uint4 s = f_limb(m, lvl, (uint32_t)(e & ((1u << logL) - 1)));
			acc ^= f_mul128_sub(T, __ldg(vec + j), s, lvl);
		}
		out[i] = acc;
	}
}

// fold_right fast path (the ring-switch shape, core/src/ring_switch/eq_ind.rs:140-146): when
// vec.len() * 2^lvl == 128 every output is a GF(2)-linear image of ONE input element,
//   out[i] = sum_j vec[j] * limb_j(mat[i]) = L(mat[i]),   L(beta_{j*2^lvl + s}) = vec[j] * beta_s,
// so the byte-LUT engine applies: 16 conflict-free LDS.128 per output instead of 128/2^lvl multiplies.
__global__ void __launch_bounds__(FOLD_THREADS, 2) k_fold_right_lut(const uint4 *__restrict__ mat, uint32_t lvl, const uint4 *__restrict__ vec,
																   uint4 *__restrict__ out, uint64_t n_out) {
	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t *tbl = smem;
	uint4 *stage = reinterpret_cast<uint4 *>(smem + LUT_BYTES);
	for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) stage[(i & 7) * 16 + (i >> 3)] = basis_image(__ldg(vec + (i >> lvl)), i & ((1u << lvl) - 1));
	__syncthreads();
	lut_build_images(tbl, stage);
	const LutLane L = lut_lane_init();
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_out; i += (uint64_t)gridDim.x * blockDim.x)
		out[i] = lut_apply(tbl, L, __ldg(mat + i));
}

// several matrices folded by the same query in one launch (the table is built once per CTA): n_out = 2^log_n_out each
struct FrSeg {
	const uint4 *mat;
	uint4 *out;
};
__global__ void __launch_bounds__(FOLD_THREADS, 2) k_fold_right_lut_multi(const FrSeg *__restrict__ segs, uint32_t n_segs, uint32_t log_n_out, uint32_t lvl,
																		 const uint4 *__restrict__ vec) {
	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t *tbl = smem;
	uint4 *stage = reinterpret_cast<uint4 *>(smem + LUT_BYTES);
	for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) stage[(i & 7) * 16 + (i >> 3)] = basis_image(__ldg(vec + (i >> lvl)), i & ((1u << lvl) - 1));
	__syncthreads();
	lut_build_images(tbl, stage);
	const LutLane L = lut_lane_init();
	const uint64_t total = (uint64_t)n_segs << log_n_out, mask = (1ull << log_n_out) - 1;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t sg = (uint32_t)(i >> log_n_out);
		const uint64_t j = i & mask;
		const uint4 *m = reinterpret_cast<const uint4 *>(__ldg(reinterpret_cast<const unsigned long long *>(&segs[sg].mat)));
		uint4 *o = reinterpret_cast<uint4 *>(__ldg(reinterpret_cast<const unsigned long long *>(&segs[sg].out)));
		o[j] = lut_apply(tbl, L, __ldg(m + j));
	}
}

// out[i] = L(in[i]) for a GF(2)-linear map L: B128 -> B128 given by its 128 basis images W[k] = L(beta_k)
// (FieldLinearTransformation::transform, field/src/linear_transformation.rs; the tower <-> POLYVAL basis change of
// convert_witnesses_to_fast_ext, core/src/constraint_system/prove.rs:291-292, with the tables of field/src/polyval.rs:
// 516-788): the byte-LUT engine, 16 conflict-free LDS.128 per element.  In place allowed (out == in).
__global__ void __launch_bounds__(FOLD_THREADS, 2) k_linear_map(const uint4 *in, const uint4 *__restrict__ W, uint4 *out, uint64_t n) {
	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t *tbl = smem;
	uint4 *stage = reinterpret_cast<uint4 *>(smem + LUT_BYTES);
	lut_build(tbl, stage, W);
	const LutLane L = lut_lane_init();
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = lut_apply(tbl, L, in[i]);
}

// fold_left fast path for B1 matrices (the switchover / evaluate_partial_high of bit-packed columns,
// math/src/fold.rs:358-516 are the reference's own B1 fast paths): with nq = vec.len() <= 128 and
// n_out a multiple of 128,
//     out[i] = sum_j vec[j] * bit(j * n_out + i)  =  L(c_i),   c_i = (bit(j * n_out + i))_j in {0,1}^128,
// L the GF(2)-linear map with basis images vec[j].  A warp takes 128 consecutive outputs: lane l loads
// the 128-bit words of rows j = l, l+32, l+64, l+96 (coalesced per row), the 128 x 128 bit block is
// transposed across the warp (sixteen 32x32 shuffle-butterfly transposes), after which every lane
// holds the vectors c_i of its 4 outputs and applies L with the byte-LUT engine (16 LDS.128 each)
// instead of 128 predicated accumulations.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, uint32_t lane) {
#pragma unroll
	for (uint32_t s = 16; s >= 1; s >>= 1) {
		const uint32_t m = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
		const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
		x = (lane & s) ? (((y >> s) & m) | (x & ~m)) : ((x & m) | ((y & m) << s));
	}
	return x;
}
__global__ void __launch_bounds__(FOLD_THREADS, 2) k_fold_left_b1_lut(const uint4 *__restrict__ mat, const uint4 *__restrict__ vec, uint32_t nq,
																	  uint4 *__restrict__ out, uint64_t n_out) {
	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t *tbl = smem;
	uint4 *stage = reinterpret_cast<uint4 *>(smem + LUT_BYTES);
	for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) stage[(i & 7) * 16 + (i >> 3)] = i < nq ? __ldg(vec + i) : u4_zero();
	__syncthreads();
	lut_build_images(tbl, stage);
	const LutLane L = lut_lane_init();
	const uint32_t lane = threadIdx.x & 31;
	const uint64_t words_per_row = n_out >> 7;  // 128-bit words per matrix row j
	const uint64_t n_blocks = words_per_row, warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	for (uint64_t blk = (((uint64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; blk < n_blocks; blk += warps) {
		uint32_t w[4][4];  // [row block r][word c]
#pragma unroll
		for (uint32_t r = 0; r < 4; r++) {
			const uint32_t j = 32 * r + lane;
			uint4 v = j < nq ? __ldg(mat + j * words_per_row + blk) : u4_zero();
			w[r][0] = v.x; w[r][1] = v.y; w[r][2] = v.z; w[r][3] = v.w;
		}
#pragma unroll
		for (uint32_t r = 0; r < 4; r++)
#pragma unroll
			for (uint32_t c = 0; c < 4; c++) w[r][c] = warp_transpose32(w[r][c], lane);
#pragma unroll
		for (uint32_t c = 0; c < 4; c++)
			out[blk * 128 + 32 * c + lane] = lut_apply(tbl, L, make_uint4(w[0][c], w[1][c], w[2][c], w[3][c]));
	}
}

// ------------------------------------------------------------------------------------------------
// ArithCircuit interpreter (math/src/arith_expr.rs:367-383)
constexpr uint32_t MAX_EXPR_STEPS = 64;
struct DevExpr {
	const b200_expr_step *steps;
	uint32_t n_steps;
	uint32_t n_vars;
};
struct PtrList {
	const uint4 *const *ptrs;  // device array of row pointers
	uint32_t n;
};

__device__ __forceinline__ uint4 f_pow128(const FieldTables &T, uint4 x, uint64_t e) {
	uint4 r = u4_one();
	while (e) {
		if (e & 1) r = f_mul128(T, r, x);
		e >>= 1;
		if (e) x = f_mul128(T, x, x);
	}
	return r;
}

__device__ __forceinline__ uint4 expr_eval(const FieldTables &T, const DevExpr &E, const uint4 *const *rows, uint64_t i) {
	uint4 tmp[MAX_EXPR_STEPS];
	for (uint32_t s = 0; s < E.n_steps; s++) {
		const b200_expr_step st = E.steps[s];
		uint4 v;
		switch (st.op) {
		case 0: v = tmp[st.l] ^ tmp[st.r]; break;
		case 1: v = f_mul128(T, tmp[st.l], tmp[st.r]); break;
		case 2: v = f_pow128(T, tmp[st.l], st.r); break;
		case 3: v = make_uint4((uint32_t)st.c_lo, (uint32_t)(st.c_lo >> 32), (uint32_t)st.c_hi, (uint32_t)(st.c_hi >> 32)); break;
		default: v = __ldg(rows[st.l] + i); break;
		}
		tmp[s] = v;
	}
	return E.n_steps ? tmp[E.n_steps - 1] : u4_zero();
}

// compute_composite (cpu/layer.rs:410-435): out[i] = expr(inputs[.][i])
__global__ void __launch_bounds__(256) k_compute_composite(const uint8_t *__restrict__ g_tables, PtrList in, DevExpr E,
														   uint4 *__restrict__ out, uint64_t n) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		out[i] = expr_eval(T, E, in.ptrs, i);
}

// KernelExecutor::sum_composition_evals (cpu/layer.rs:499-514): slot += coeff * sum_i expr(rows[.][i])
__global__ void __launch_bounds__(256) k_sum_composition(const uint8_t *__restrict__ g_tables, PtrList in, DevExpr E, uint64_t n,
														 uint4 coeff, uint4 *__restrict__ slot) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	__shared__ uint4 red[32];
	uint4 acc = u4_zero();
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		acc ^= expr_eval(T, E, in.ptrs, i);
	acc = block_xor(acc, red);
	if (threadIdx.x == 0 && !is_zero(acc)) atomic_xor_u4(slot, f_mul128(T, acc, coeff));
}

// pairwise_product_reduce, one round (cpu/layer.rs:437-484): out[i] = in[2i] * in[2i+1]
__global__ void __launch_bounds__(256) k_pairwise_product(const uint8_t *__restrict__ g_tables, const uint4 *__restrict__ in,
														  uint4 *__restrict__ out, uint64_t n_out) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_out; i += (uint64_t)gridDim.x * blockDim.x)
		out[i] = f_mul128(T, __ldg(in + 2 * i), __ldg(in + 2 * i + 1));
}

// ------------------------------------------------------------------------------------------------
// v3::calculate_round_evals (core/src/protocols/sumcheck/v3/bivariate_product.rs:303-408), fused:
// blockIdx.y = composition c; slots[0] ^= alpha^c * sum_i hi_a*hi_b ; slots[1] ^= alpha^c * sum_i
// (lo_a+hi_a)*(lo_b+hi_b).  `pows[c]` = alpha^c.
__global__ void __launch_bounds__(256) k_bivariate_round_evals(const uint8_t *__restrict__ g_tables, const uint4 *const *__restrict__ mls,
															   uint64_t half, const uint32_t *__restrict__ ia, const uint32_t *__restrict__ ib,
															   const uint4 *__restrict__ pows, uint4 *__restrict__ slot_y1,
															   uint4 *__restrict__ slot_yinf) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	__shared__ uint4 red[32];
	uint32_t c = blockIdx.y;
	const uint4 *a = mls[ia[c]], *b = mls[ib[c]];
	uint4 s1 = u4_zero(), sinf = u4_zero();
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < half; i += (uint64_t)gridDim.x * blockDim.x) {
		uint4 alo = __ldg(a + i), ahi = __ldg(a + half + i), blo = __ldg(b + i), bhi = __ldg(b + half + i);
		s1 ^= f_mul128(T, ahi, bhi);
		sinf ^= f_mul128(T, alo ^ ahi, blo ^ bhi);
	}
	s1 = block_xor(s1, red);
	sinf = block_xor(sinf, red);
	if (threadIdx.x == 0) {
		uint4 p = pows[c];
		if (!is_zero(s1)) atomic_xor_u4(slot_y1, f_mul128(T, s1, p));
		if (!is_zero(sinf)) atomic_xor_u4(slot_yinf, f_mul128(T, sinf, p));
	}
}

// ------------------------------------------------------------------------------------------------
// eq-ind (zerocheck) round evaluations, HighToLow: R[c][p] = sum_i E[i] * C_c^{(p)}(P(i))
// reference: hal/src/sumcheck_round_calculation.rs:222-297, core/.../prove/eq_ind.rs:670-702.
// point code 1: P = hi ; 2: P = hi - lo evaluated on leading_term(C) ; >=3: P = lo + (hi-lo)*z_p.
// blockIdx.y = c * n_points + p
struct EqIndArgs {
	const uint4 *const *mls;
	const uint64_t *lens;   // [n_mls] stored prefix length of each multilinear (<= 2*half)
	const uint4 *suffix;    // [n_mls] implicit constant value beyond the stored prefix
	uint32_t n_mls;
	uint64_t half;
	const uint4 *eq_ind;
	const DevExpr *comps;       // [n_comp]
	const DevExpr *comps_lead;  // [n_comp]
	const uint32_t *codes;      // [n_points]
	const uint4 *points;        // [n_points]
	uint32_t n_points;
	uint4 *slots;  // [n_comp * n_points]
	uint32_t low_to_high;  // 0: HighToLow pairs (i, half + i); 1: LowToHigh pairs (2i, 2i + 1)
};
constexpr uint32_t MAX_EQIND_MLS = 24;  // per-thread operand staging for the general-point path

// C^{(p)}(P(i)) for one composition / evaluation point / hypercube index
// NC: operands through the read-only path (__ldg); false inside the persistent tail kernel, which folds the multilinears
// in place between rounds (the non-coherent path could return a stale line)
template <bool NC = true>
__device__ __forceinline__ uint4 eq_ind_eval_point(const FieldTables &T, const EqIndArgs &A, const DevExpr &E, uint32_t code, uint4 z, uint64_t i) {
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
			const uint4 *m = A.mls[st.l];
			const uint64_t len = A.lens[st.l];
			// HighToLowAccess / LowToHighAccess (sumcheck_round_calculation.rs:408-604)
			const uint64_t i_lo = A.low_to_high ? 2 * i : i, i_hi = A.low_to_high ? 2 * i + 1 : A.half + i;
			uint4 hi = i_hi < len ? (NC ? __ldg(m + i_hi) : m[i_hi]) : A.suffix[st.l];
			if (code == 1) v = hi;
			else {
				uint4 lo = i_lo < len ? (NC ? __ldg(m + i_lo) : m[i_lo]) : A.suffix[st.l];
				uint4 d = hi ^ lo;
				v = code == 2 ? d : (lo ^ f_mul128(T, z, d));  // warp-uniform operand FIRST (field.cuh)
			}
		}
		}
		tmp[s] = v;
	}
	return E.n_steps ? tmp[E.n_steps - 1] : u4_zero();
}

__global__ void __launch_bounds__(256) k_eq_ind_round_evals(const uint8_t *__restrict__ g_tables, EqIndArgs A) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	__shared__ uint4 red[32];
	uint32_t c = blockIdx.y / A.n_points, p = blockIdx.y % A.n_points;
	uint32_t code = A.codes[p];
	const DevExpr E = code == 2 ? A.comps_lead[c] : A.comps[c];
	uint4 z = A.points[p];
	uint4 acc = u4_zero();
	// eq_ind == nullptr: the regular (unweighted) evaluator, prove/regular_sumcheck.rs:233-277
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < A.half; i += (uint64_t)gridDim.x * blockDim.x) {
		uint4 v = eq_ind_eval_point(T, A, E, code, z, i);
		acc ^= A.eq_ind ? f_mul128(T, v, __ldg(A.eq_ind + i)) : v;
	}
	acc = block_xor(acc, red);
	if (threadIdx.x == 0) atomic_xor_u4(A.slots + blockIdx.y, acc);
}

// Large rounds: the composition values are materialised (vals[(c*n_points + p) * half + i]) and the
// weighted sums sum_i E[i] * val[i] run as inner-product jobs on the tensor cores (roundevals_tc.cuh),
// which removes the per-point multiplication by the eq-indicator from the ALU path.
__global__ void __launch_bounds__(256) k_eq_ind_vals(const uint8_t *__restrict__ g_tables, EqIndArgs A, uint4 *__restrict__ vals) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	uint32_t c = blockIdx.y / A.n_points, p = blockIdx.y % A.n_points;
	uint32_t code = A.codes[p];
	const DevExpr E = code == 2 ? A.comps_lead[c] : A.comps[c];
	uint4 z = A.points[p];
	uint4 *out = vals + (uint64_t)blockIdx.y * A.half;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < A.half; i += (uint64_t)gridDim.x * blockDim.x)
		out[i] = eq_ind_eval_point(T, A, E, code, z, i);
}

// Degree-2 monomials of the monomial plan (eqind_plan.hpp): w[i] = E[i] * (x0[i] ^ x1[i]) (x1 may be
// null).  blockIdx.y = scaled vector.
struct EqScaleOp {
	const uint4 *x0, *x1;
	uint4 *w;
};
__global__ void __launch_bounds__(256) k_eq_scale(const uint8_t *__restrict__ g_tables, const uint4 *__restrict__ eq_ind,
												  const EqScaleOp *__restrict__ ops, uint64_t len) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	const EqScaleOp op = ops[blockIdx.y];
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x) {
		uint4 x = __ldg(op.x0 + i);
		if (op.x1) x = x ^ __ldg(op.x1 + i);
		op.w[i] = f_mul128(T, x, __ldg(eq_ind + i));
	}
}

// ------------------------------------------------------------------------------------------------
// fri_fold (compute/src/cpu/layer.rs:304-391).  One thread per output element, chunk of
// 2^n_ch <= 2^FRI_MAX_LOG_CHUNK values kept in local memory.  Twiddles (field T_kt, kt <= 5) are
// produced on the fly from the s_evals rows: t = XOR_b bit_b(idx) * s[row][b]  (twiddle.rs:163-168).
constexpr uint32_t FRI_MAX_LOG_CHUNK = 8;
struct FriArgs {
	const uint4 *in;
	uint4 *out;
	uint64_t n_out;
	uint32_t log_len, log_batch, n_ch;
	const uint4 *challenges;  // device [n_ch]
	const uint32_t *s_evals;  // device [d][32] rows (padded), twiddle field elements (<= 32 bit)
	uint32_t d;               // log domain size
	uint32_t kt;              // twiddle field level
};

__device__ __forceinline__ uint32_t twiddle_on_the_fly(const uint32_t *s_row, uint32_t n_bits, uint64_t idx) {
	uint32_t t = 0;
	for (uint32_t b = 0; b < n_bits; b++)
		if ((idx >> b) & 1) t ^= s_row[b];
	return t;
}

__global__ void __launch_bounds__(128) k_fri_fold(const uint8_t *__restrict__ g_tables, FriArgs A) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	uint32_t eta = A.n_ch - A.log_batch;
	uint32_t chunk = 1u << A.n_ch;
	for (uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < A.n_out; c += (uint64_t)gridDim.x * blockDim.x) {
		uint4 v[1u << FRI_MAX_LOG_CHUNK];
		for (uint32_t i = 0; i < chunk; i++) v[i] = __ldg(A.in + c * chunk + i);
		uint32_t cur = chunk;
		for (uint32_t r = 0; r < A.log_batch; r++) {
			cur >>= 1;
			uint4 z = A.challenges[r];
			for (uint32_t o = 0; o < cur; o++) v[o] = v[2 * o] ^ f_mul128(T, z, v[2 * o] ^ v[2 * o + 1]);
		}
		uint32_t L = A.log_len, sz = eta;
		for (uint32_t r = 0; r < eta; r++) {
			uint4 z = A.challenges[A.log_batch + r];
			uint32_t row = A.d - L;
			for (uint32_t o = 0; o < (1u << (sz - 1)); o++) {
				uint32_t t = twiddle_on_the_fly(A.s_evals + row * 32, A.d - 1 - row, (c << (sz - 1)) | o);
				uint4 u = v[2 * o], w = v[2 * o + 1];
				w ^= u;
				u ^= f_mul128_sub(T, w, make_uint4(t, 0, 0, 0), A.kt);
				v[o] = u ^ f_mul128(T, z, u ^ w);
			}
			L--;
			sz--;
		}
		A.out[c] = v[0];
	}
}

// Fast path for n_ch = NCH <= 5 challenges (the prover's arities): the chunk lives in registers and
// every lerp `u + (v-u)*challenge` is a multiplication by one of NCH broadcast constants, served by
// the nibble-LUT engine (8 KiB of shared memory per challenge) instead of the general multiply.
// dyn smem = FIELD_TABLE_BYTES + NCH * NLUT_BYTES
template <uint32_t NCH>
__global__ void __launch_bounds__(256) k_fri_fold_lut(const uint8_t *__restrict__ g_tables, FriArgs A) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	uint8_t *luts = smem + FIELD_TABLE_BYTES;
	for (uint32_t r = 0; r < NCH; r++) nlut_build_mul(luts + r * NLUT_BYTES, A.challenges[r]);
	const NLutLane L = nlut_lane_init();
	constexpr uint32_t CHUNK = 1u << NCH;
	const uint32_t log_batch = A.log_batch;
	for (uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < A.n_out; c += (uint64_t)gridDim.x * blockDim.x) {
		uint4 v[CHUNK];
#pragma unroll
		for (uint32_t i = 0; i < CHUNK; i++) v[i] = __ldg(A.in + c * CHUNK + i);
#pragma unroll
		for (uint32_t r = 0; r < NCH; r++) {
			const uint8_t *lut = luts + r * NLUT_BYTES;
			if (r < log_batch) {
#pragma unroll
				for (uint32_t o = 0; o < (CHUNK >> (r + 1)); o++) v[o] = v[2 * o] ^ nlut_apply(lut, L, v[2 * o] ^ v[2 * o + 1]);
			} else {
				const uint32_t k = r - log_batch;             // fold round index
				const uint32_t Llen = A.log_len - k, sz = (NCH - log_batch) - k;
				const uint32_t row = A.d - Llen;
#pragma unroll
				for (uint32_t o = 0; o < (CHUNK >> (r + 1)); o++) {
					uint32_t t = twiddle_on_the_fly(A.s_evals + row * 32, A.d - 1 - row, (c << (sz - 1)) | o);
					uint4 u = v[2 * o], w = v[2 * o + 1];
					w ^= u;
					u ^= f_mul128_sub(T, w, make_uint4(t, 0, 0, 0), A.kt);
					v[o] = u ^ nlut_apply(lut, L, u ^ w);
				}
			}
		}
		A.out[c] = v[0];
	}
}

// fri_fold with eta = 0 (all NCH challenges are tensor-lerp rounds: the first fold of the interleaved
// codeword, fri/prove.rs:343-386 with log_batch = arity): per output a binary lerp tree over 2^NCH
// consecutive inputs.  Rounds 0..2 (8 + 4 + 2 of the 15 products at NCH = 4) use one Karatsuba-64 table
// each (linmap.cuh), a fourth round the 8 KiB nibble table.  dyn smem = min(NCH,3)*LUT_BYTES + 6144 + NLUT_BYTES
//
// A warp owns 32 consecutive outputs = 32 * 2^NCH consecutive inputs and reads them COALESCED: lane l
// loads elements j*32 + l.  The lerp tree then runs across lanes: in round r the two children of a node
// sit in lanes that differ in bit r, so lanes exchange one value per product with __shfl_xor (the lane
// with bit r = 0 finishes the node of register 2q, its partner the node of register 2q+1) and every lane
// does the same number of products in every round.  After NCH rounds lane l holds one output.
constexpr uint32_t FRI_K64_THREADS = 512;
template <uint32_t NCH>
__global__ void __launch_bounds__(FRI_K64_THREADS, 1) k_fri_lerp_k64(FriArgs A) {
	extern __shared__ __align__(256) uint8_t smem[];
	constexpr uint32_t NK = NCH < 3 ? NCH : 3;
	uint2 *stage = reinterpret_cast<uint2 *>(smem + NK * LUT_BYTES);
	uint8_t *nl = smem + NK * LUT_BYTES + 6144;
	uint4 zs[NK];
#pragma unroll
	for (uint32_t t = 0; t < NK; t++) zs[t] = A.challenges[t];
	k64_build_multi<NK>(smem, stage, zs);
	if (NCH > 3) nlut_build_mul(nl, A.challenges[3]);
	K64Lane L = k64_lane_init(smem);
	const uint32_t sbase = L.sbase;
	const NLutLane NL = nlut_lane_init();
	constexpr uint32_t CHUNK = 1u << NCH;
	const uint32_t lane = threadIdx.x & 31;
	// output index of this lane inside its warp-tile
	// the lane whose bit r is set finishes the odd register of round r, so it ends up with register
	// j = lane mod 2^NCH of the load phase, i.e. output j * (32 >> NCH) + (lane >> NCH)
	const uint32_t o_lane = (lane & (CHUNK - 1)) * (32u >> NCH) + (lane >> NCH);
	const uint64_t n_tiles = A.n_out >> 5, warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;  // n_out is a multiple of 32
	for (uint64_t tile = (((uint64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; tile < n_tiles; tile += warps) {
		const uint4 *src = A.in + tile * (32 * CHUNK) + lane;
		uint4 v[CHUNK];
#pragma unroll
		for (uint32_t j = 0; j < CHUNK; j++) v[j] = __ldg(src + 32 * j);
#pragma unroll
		for (uint32_t r = 0; r < NCH; r++) {
			if (r < NK) L.sbase = sbase + r * LUT_BYTES;
			const bool odd = (lane >> r) & 1u;
#pragma unroll
			for (uint32_t q = 0; q < (CHUNK >> (r + 1)); q++) {
				// even-side lanes finish register 2q (they hold child 0), odd-side lanes register 2q+1 (child 1)
				const uint4 mine = odd ? v[2 * q + 1] : v[2 * q], send = odd ? v[2 * q] : v[2 * q + 1];
				uint4 recv;
				recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1u << r);
				recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1u << r);
				recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1u << r);
				recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1u << r);
				const uint4 x = mine ^ recv;          // child 0 + child 1
				const uint4 c0 = odd ? recv : mine;   // child 0
				v[q] = c0 ^ (r < NK ? k64_apply(L, x) : nlut_apply(nl, NL, x));
			}
		}
		A.out[tile * 32 + o_lane] = v[0];
	}
}

}  // namespace b200
