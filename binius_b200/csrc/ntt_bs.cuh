// ntt_bs.cuh -- bit-sliced additive-NTT pass kernel for B32 (the fast path).
//
// B200 has no GF(2^k) multiplier (no GFNI / CLMUL) and a table-driven B32 multiply costs ~13
// data-dependent shared-memory gathers, so the butterflies are computed BIT-SLICED: 32 scalars that
// share every twiddle (32 consecutive positions of the contiguous "inner" index x | y_low << log_x)
// are transposed into 32 words (word p = bit p of the 32 scalars) and multiplied by the twiddle with
// a register-resident Karatsuba circuit over the tower (T(32) ~ 1.1k LOP3 per 32 products, i.e.
// ~34 ALU ops per B32 product instead of ~120 instructions + 13 LDS).
//
// Reference semantics: crates/ntt/src/tests/reference.rs:68-160, crates/ntt/src/single_threaded.rs:
// 134-362 (see ntt.cuh).  Tower multiply: pairwise_recursive_arithmetic.rs:12-62 applied bit-plane-wise.
#pragma once
#include "field.cuh"

namespace b200 {

// ---- 32x32 bit-matrix transpose held in registers (a[i] = row i) ---------------------------------
__device__ __forceinline__ void transpose32(uint32_t (&a)[32]) {
	// stages 16 and 8 move whole half-words / bytes: one PRMT per output word
#pragma unroll
	for (int k = 0; k < 16; k++) {
		const uint32_t lo = a[k], hi = a[k + 16];
		a[k] = __byte_perm(lo, hi, 0x5410);
		a[k + 16] = __byte_perm(lo, hi, 0x7632);
	}
#pragma unroll
	for (int k = 0; k < 32; k++) {
		if ((k & 8) == 0) {
			const uint32_t lo = a[k], hi = a[k + 8];
			a[k] = __byte_perm(lo, hi, 0x6240);
			a[k + 8] = __byte_perm(lo, hi, 0x7351);
		}
	}
#pragma unroll
	for (int j = 4; j >= 1; j >>= 1) {
		const uint32_t m = j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
		for (int k = 0; k < 32; k++) {
			if ((k & j) == 0) {
				uint32_t t = ((a[k] >> j) ^ a[k + j]) & m;
				a[k + j] ^= t;
				a[k] ^= t << j;
			}
		}
	}
}

// ---- bit-sliced tower arithmetic: an element of T_k is N = 2^k words (word p = bit p) ------------
template <int N> struct BS {
	static constexpr int H = N / 2;
	// r = a * X_{k-1}  (mul_alpha at level k): (a0, a1) -> (a1, a0 + alpha(a1))
	static __device__ __forceinline__ void alpha(const uint32_t (&a)[N], uint32_t (&r)[N]) {
		uint32_t a1[H], t[H];
#pragma unroll
		for (int i = 0; i < H; i++) a1[i] = a[H + i];
		BS<H>::alpha(a1, t);
#pragma unroll
		for (int i = 0; i < H; i++) {
			r[i] = a1[i];
			r[H + i] = a[i] ^ t[i];
		}
	}
	// r = a * b
	static __device__ __forceinline__ void mul(const uint32_t (&a)[N], const uint32_t (&b)[N], uint32_t (&r)[N]) {
		uint32_t a0[H], a1[H], b0[H], b1[H], sa[H], sb[H], z0[H], z2[H], m[H], al[H];
#pragma unroll
		for (int i = 0; i < H; i++) {
			a0[i] = a[i]; a1[i] = a[H + i];
			b0[i] = b[i]; b1[i] = b[H + i];
			sa[i] = a0[i] ^ a1[i];
			sb[i] = b0[i] ^ b1[i];
		}
		BS<H>::mul(a0, b0, z0);
		BS<H>::mul(a1, b1, z2);
		BS<H>::mul(sa, sb, m);
		BS<H>::alpha(z2, al);
#pragma unroll
		for (int i = 0; i < H; i++) {
			uint32_t lo = z0[i] ^ z2[i];
			r[i] = lo;
			r[H + i] = m[i] ^ lo ^ al[i];
		}
	}
};
// GF(4) = T_1: (a0 + a1 X)(b0 + b1 X), X^2 = X + 1:  lo = a0b0 ^ a1b1, hi = a0b1 ^ a1b0 ^ a1b1.
// Written as and-xor chains so that every step is ONE LOP3 (4 ops instead of the 6 of the Karatsuba form).
template <> struct BS<2> {
	static __device__ __forceinline__ void alpha(const uint32_t (&a)[2], uint32_t (&r)[2]) {
		r[0] = a[1];
		r[1] = a[0] ^ a[1];
	}
	static __device__ __forceinline__ void mul(const uint32_t (&a)[2], const uint32_t (&b)[2], uint32_t (&r)[2]) {
		uint32_t t = a[1] & b[1];
		r[0] = (a[0] & b[0]) ^ t;
		uint32_t u = (a[0] & b[1]) ^ t;
		r[1] = (a[1] & b[0]) ^ u;
	}
};
template <> struct BS<1> {
	static __device__ __forceinline__ void alpha(const uint32_t (&a)[1], uint32_t (&r)[1]) { r[0] = a[0]; }
	static __device__ __forceinline__ void mul(const uint32_t (&a)[1], const uint32_t (&b)[1], uint32_t (&r)[1]) { r[0] = a[0] & b[0]; }
};

// The same circuit for a SCALAR second operand: b[i] in {0, 1} (bit i of a twiddle shared by the 32
// bit-sliced lanes).  A leaf product a & (0 - b) is then the integer product a * b, which issues on the
// FMA pipe (IMAD) instead of the ALU pipe that carries all the XORs: the kernel is ALU-pipe bound, and
// this moves the 324 leaf ANDs of a T(32) off it (2 LOP3 + 4 IMAD per GF(4) product instead of 4 LOP3).
template <int N> struct BSs {
	static constexpr int H = N / 2;
	static __device__ __forceinline__ void mul(const uint32_t (&a)[N], const uint32_t (&b)[N], uint32_t (&r)[N]) {
		uint32_t a0[H], a1[H], b0[H], b1[H], sa[H], sb[H], z0[H], z2[H], m[H], al[H];
#pragma unroll
		for (int i = 0; i < H; i++) {
			a0[i] = a[i]; a1[i] = a[H + i];
			b0[i] = b[i]; b1[i] = b[H + i];
			sa[i] = a0[i] ^ a1[i];
			sb[i] = b0[i] ^ b1[i];
		}
		BSs<H>::mul(a0, b0, z0);
		BSs<H>::mul(a1, b1, z2);
		BSs<H>::mul(sa, sb, m);
		BS<H>::alpha(z2, al);
#pragma unroll
		for (int i = 0; i < H; i++) {
			uint32_t lo = z0[i] ^ z2[i];
			r[i] = lo;
			r[H + i] = m[i] ^ lo ^ al[i];
		}
	}
};
template <> struct BSs<2> {
	static __device__ __forceinline__ void mul(const uint32_t (&a)[2], const uint32_t (&b)[2], uint32_t (&r)[2]) {
		const uint32_t p00 = a[0] * b[0], p11 = a[1] * b[1], p01 = a[0] * b[1], p10 = a[1] * b[0];
		r[0] = p00 ^ p11;
		r[1] = p01 ^ p10 ^ p11;
	}
};
template <> struct BSs<1> {
	static __device__ __forceinline__ void mul(const uint32_t (&a)[1], const uint32_t (&b)[1], uint32_t (&r)[1]) { r[0] = a[0] * b[0]; }
};

// v (bit-sliced, 32 planes) times a scalar twiddle t: p = v * t.  Sub-field twiddles use the
// limb-wise product (binary_field.rs:363-393): B32 x B16 = 2 x T(16), B32 x B8 = 4 x T(8).
// `bits` = the twiddle expanded to one byte (0 / 1) per bit in shared memory: the operand words come
// straight out of LDS.U8 (the LSU pipe idles in this kernel) instead of 2 ALU ops per bit.
template <int N>
__device__ __forceinline__ void bs_mul_limbs(const uint32_t (&v)[32], const uint8_t *bits, uint32_t (&p)[32]) {
	uint32_t b[N];
#pragma unroll
	for (int i = 0; i < N; i++) b[i] = bits[i];
#pragma unroll
	for (int l = 0; l < 32 / N; l++) {
		uint32_t a[N], r[N];
#pragma unroll
		for (int i = 0; i < N; i++) a[i] = v[l * N + i];
		BSs<N>::mul(a, b, r);
#pragma unroll
		for (int i = 0; i < N; i++) p[l * N + i] = r[i];
	}
}

__device__ __forceinline__ void bs_mul_scalar(const uint32_t (&v)[32], uint32_t t, const uint8_t *bits, uint32_t (&p)[32]) {
	if (t >> 16) bs_mul_limbs<32>(v, bits, p);
	else if (t >> 8) bs_mul_limbs<16>(v, bits, p);
	else if (t >> 1) bs_mul_limbs<8>(v, bits, p);
	else {
		uint32_t m = 0u - (t & 1u);
#pragma unroll
		for (int i = 0; i < 32; i++) p[i] = v[i] & m;
	}
}

// t -> 32 bytes of 0 / 1 (8 words: nibble q of t spread over the bytes of word q)
__device__ __forceinline__ void ntt_store_bits(uint8_t *dst, uint32_t t) {
	uint4 lo, hi;
	auto spread = [](uint32_t nib) { return ((nib & 0xFu) * 0x00204081u) & 0x01010101u; };
	lo = make_uint4(spread(t), spread(t >> 4), spread(t >> 8), spread(t >> 12));
	hi = make_uint4(spread(t >> 16), spread(t >> 20), spread(t >> 24), spread(t >> 28));
	reinterpret_cast<uint4 *>(dst)[0] = lo;
	reinterpret_cast<uint4 *>(dst)[1] = hi;
}

struct NttBsArgs {
	uint32_t *data;
	uint32_t log_x, log_y;  // log_x includes the extension-degree shift
	uint32_t i_lo, R;       // layers [i_lo, i_lo + R)
	uint32_t log_cu;        // log2(units per row); a unit = 32 consecutive inner positions
	uint32_t row0, d;
	uint64_t coset;
	int inverse;
	const uint32_t *s_evals;
	// between two bit-sliced passes the data stays bit-sliced in HBM (every 128-byte unit holds its 32
	// bit-planes instead of its 32 scalars): only the first pass transposes in and the last one out
	uint32_t in_sliced, out_sliced;
};

// Tiles of 2^LT units, LT = 8 or 9 (32 / 64 KiB of planes), one butterfly-unit per thread and layer: 128
// threads x 4 CTAs per SM or 256 threads x 2 CTAs per SM.  Both fit the register file at 128 registers per
// thread (ptxas fits the 32-plane product there with ~150 bytes of spills; 168 registers are spill-free and
// as fast).  A pass has many more CTAs than CTA slots, so the tail of the grid is short (1024-unit tiles on
// one CTA per SM ran a 3.46-wave grid at 2^24), the global load/store phases of one CTA overlap the
// butterflies of the others, and the per-layer barriers are decoupled.  The host picks LT per transform:
// 9 up to 2^25 coefficients (fewer passes: 17 layers = 8 + 9), 8 above (measured: RS-encode shape at 2^24
// 0.348 vs 0.358 ms, at 2^30 29.6 vs 29.1 ms).
constexpr uint32_t NTT_BS_MAX_THREADS = 256;
constexpr uint32_t NTT_BS_MAX_LOG_TILE = 9;
__host__ __device__ constexpr uint32_t ntt_bs_threads(uint32_t log_tile) { return 128u << (log_tile - 8); }

// dyn smem = 36 * 2^R (twiddles + their bit expansion) + 128 * 2^(R + log_cu) (bit-sliced tile, plane-major: tile[p * NU + unit])
__global__ void __launch_bounds__(NTT_BS_MAX_THREADS, 2) k_ntt_bs_pass(const NttBsArgs A) {
	extern __shared__ __align__(128) uint8_t smem[];
	uint32_t *tw = reinterpret_cast<uint32_t *>(smem);
	uint8_t *twb = smem + (((4u << A.R) + 15u) & ~15u);  // [2^R][32] bit-expanded twiddles
	uint32_t *tile = reinterpret_cast<uint32_t *>(twb + (32u << A.R));
	const uint32_t R = A.R, i_hi = A.i_lo + R, log_cu = A.log_cu;
	const uint32_t log_inner = A.log_x + A.i_lo;           // >= 5
	const uint32_t log_chunks = log_inner - 5 - log_cu;    // chunks of 2^log_cu units along the inner axis
	const uint32_t NU = 1u << (R + log_cu);
	uint64_t bid = blockIdx.x;
	uint64_t chunk = bid & ((1ull << log_chunks) - 1);
	uint64_t outer = bid >> log_chunks;
	uint32_t *base = A.data + ((uint64_t)blockIdx.y << (A.log_x + A.log_y)) + (outer << (i_hi + A.log_x)) + (chunk << (5 + log_cu));

	for (uint32_t e = threadIdx.x + 1; e < (1u << R); e += blockDim.x) {
		uint32_t lvl_bits = 31 - __clz(e);
		uint32_t li = R - 1 - lvl_bits;
		uint32_t jr = e - (1u << lvl_bits);
		uint32_t i = A.i_lo + li;
		uint64_t idx = (A.coset << (A.log_y - 1 - i)) | (outer << (i_hi - i - 1)) | jr;
		uint32_t row = A.row0 + i;
		uint32_t nb = A.d - 1 - row;
		const uint32_t *srow = A.s_evals + row * 32;
		uint32_t t = 0;
		for (uint32_t b = 0; b < nb; b++)
			if ((idx >> b) & 1) t ^= srow[b];
		tw[e] = t;
		ntt_store_bits(twb + e * 32, t);
	}
	// load + transpose
	for (uint32_t uid = threadIdx.x; uid < NU; uid += blockDim.x) {
		uint32_t r = uid >> log_cu, c = uid & ((1u << log_cu) - 1);
		const uint4 *src = reinterpret_cast<const uint4 *>(base + ((uint64_t)r << log_inner) + c * 32);
		uint32_t a[32];
#pragma unroll
		for (int q = 0; q < 8; q++) {
			uint4 w = src[q];
			a[4 * q] = w.x; a[4 * q + 1] = w.y; a[4 * q + 2] = w.z; a[4 * q + 3] = w.w;
		}
		if (!A.in_sliced) transpose32(a);
#pragma unroll
		for (int p = 0; p < 32; p++) tile[p * NU + uid] = a[p];
	}
	__syncthreads();
	const uint32_t n_bf = NU >> 1;
	for (uint32_t step = 0; step < R; step++) {
		uint32_t li = A.inverse ? step : (R - 1 - step);
		for (uint32_t bfi = threadIdx.x; bfi < n_bf; bfi += blockDim.x) {
			uint32_t c = bfi & ((1u << log_cu) - 1), q = bfi >> log_cu;
			uint32_t kk = q & ((1u << li) - 1), jr = q >> li;
			uint32_t r0 = (jr << (li + 1)) | kk, r1 = r0 | (1u << li);
			const uint32_t te = (1u << (R - 1 - li)) + jr;
			const uint32_t t = tw[te];
			const uint8_t *tb = twb + te * 32;
			uint32_t iu = (r0 << log_cu) | c, iv = (r1 << log_cu) | c;
			uint32_t v[32], p[32];
#pragma unroll
			for (int k = 0; k < 32; k++) v[k] = tile[k * NU + iv];
			if (!A.inverse) {
				bs_mul_scalar(v, t, tb, p);
#pragma unroll
				for (int k = 0; k < 32; k++) {
					uint32_t u = tile[k * NU + iu] ^ p[k];
					tile[k * NU + iu] = u;
					tile[k * NU + iv] = v[k] ^ u;
				}
			} else {
#pragma unroll
				for (int k = 0; k < 32; k++) {
					v[k] ^= tile[k * NU + iu];
					tile[k * NU + iv] = v[k];
				}
				bs_mul_scalar(v, t, tb, p);
#pragma unroll
				for (int k = 0; k < 32; k++) tile[k * NU + iu] ^= p[k];
			}
		}
		__syncthreads();
	}
	for (uint32_t uid = threadIdx.x; uid < NU; uid += blockDim.x) {
		uint32_t r = uid >> log_cu, c = uid & ((1u << log_cu) - 1);
		uint4 *dst = reinterpret_cast<uint4 *>(base + ((uint64_t)r << log_inner) + c * 32);
		uint32_t a[32];
#pragma unroll
		for (int p = 0; p < 32; p++) a[p] = tile[p * NU + uid];
		if (!A.out_sliced) transpose32(a);
#pragma unroll
		for (int q = 0; q < 8; q++) dst[q] = make_uint4(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
	}
}


// ------------------------------------------------------------------------------------------------
// Lowest pass when log_x < 5: a unit is 32 CONSECUTIVE scalars (x and the low L0 = 5 - log_x bits of
// y), so layers i < L0 pair lanes INSIDE a unit and every butterfly of such a layer may have its own
// twiddle  t(j) = t_unit(y >> L0) ^ t_lane(y_low >> (i+1))  (twiddles are GF(2)-linear in the block
// index: twiddle.rs:163-168).  They are executed on pairs of units (A, B): the v-lanes of A are
// shifted onto its u-lanes, the v-lanes of B stay in place, so all 32 lanes of the general
// bit-sliced multiply do useful work.  Layers i >= L0 pair whole units as in k_ntt_bs_pass.
struct NttBsLowArgs {
	uint32_t *data;
	uint32_t log_x, log_y;
	uint32_t Rt;        // tile = 2^Rt contiguous units
	uint32_t n_intra;   // executed intra-unit layers  [0, n_intra),  n_intra <= L0 = 5 - log_x
	uint32_t n_inter;   // executed inter-unit layers  [L0, L0 + n_inter), n_inter <= Rt
	uint32_t row0, d;
	uint64_t coset;
	int inverse;
	const uint32_t *s_evals;
	uint32_t in_sliced, out_sliced;  // as in NttBsArgs
};

__device__ __forceinline__ uint32_t ntt_subset_sum(const uint32_t *srow, uint32_t nb, uint64_t idx) {
	uint32_t t = 0;
	for (uint32_t b = 0; b < nb && (idx >> b); b++)
		if ((idx >> b) & 1) t ^= srow[b];
	return t;
}

// dyn smem = 4*2^Rt (inter twiddles) + 5*4*2^Rt (per-unit intra twiddles) + 5*32*4 (lane planes) + 32*2^Rt (bit-expanded twiddles) + 128*2^Rt (tile)
__global__ void __launch_bounds__(NTT_BS_MAX_THREADS, 2) k_ntt_bs_low(const NttBsLowArgs A) {
	extern __shared__ __align__(128) uint8_t smem[];
	const uint32_t Rt = A.Rt, NU = 1u << Rt, L0 = 5 - A.log_x;
	uint32_t *tw = reinterpret_cast<uint32_t *>(smem);            // [2^Rt] heap for inter-unit layers
	uint32_t *thi = tw + NU;                                       // [5][2^Rt]
	uint32_t *tl = thi + 5 * NU;                                   // [5][32] bit-sliced lane twiddles
	uint8_t *twb = smem + ((24u * NU + 640u + 15u) & ~15u);        // [NU][32] bit-expanded inter-unit twiddles
	uint32_t *tile = reinterpret_cast<uint32_t *>(twb + 32 * NU);  // [32][NU]
	const uint64_t outer = blockIdx.x;                             // unit-index bits above the tile
	uint32_t *base = A.data + ((uint64_t)blockIdx.y << (A.log_x + A.log_y)) + (outer << (Rt + 5));

	for (uint32_t e = threadIdx.x + 1; e < NU; e += blockDim.x) {
		uint32_t lvl_bits = 31 - __clz(e);
		uint32_t li = Rt - 1 - lvl_bits;
		if (li >= A.n_inter) { tw[e] = 0; ntt_store_bits(twb + e * 32, 0); continue; }
		uint32_t jr = e - (1u << lvl_bits);
		uint32_t i = L0 + li;
		uint64_t idx = (A.coset << (A.log_y - 1 - i)) | (outer << (Rt - li - 1)) | jr;
		uint32_t row = A.row0 + i;
		tw[e] = ntt_subset_sum(A.s_evals + row * 32, A.d - 1 - row, idx);
		ntt_store_bits(twb + e * 32, tw[e]);
	}
	for (uint32_t e = threadIdx.x; e < A.n_intra * NU; e += blockDim.x) {
		uint32_t i = e >> Rt, r = e & (NU - 1);
		uint64_t yhi = (outer << Rt) | r;
		uint64_t idx = (A.coset << (A.log_y - 1 - i)) | (yhi << (L0 - 1 - i));
		uint32_t row = A.row0 + i;
		thi[i * NU + r] = ntt_subset_sum(A.s_evals + row * 32, A.d - 1 - row, idx);
	}
	if (threadIdx.x < 32) {
		uint32_t lane = threadIdx.x;
		for (uint32_t i = 0; i < A.n_intra; i++) {
			uint32_t row = A.row0 + i;
			uint32_t t = ntt_subset_sum(A.s_evals + row * 32, A.d - 1 - row, (uint64_t)((lane >> A.log_x) >> (i + 1)));
			for (uint32_t p = 0; p < 32; p++) {
				uint32_t w = __ballot_sync(0xffffffffu, (t >> p) & 1u);
				if (lane == 0) tl[i * 32 + p] = w;
			}
		}
	}
	for (uint32_t uid = threadIdx.x; uid < NU; uid += blockDim.x) {
		const uint4 *src = reinterpret_cast<const uint4 *>(base + (uint64_t)uid * 32);
		uint32_t a[32];
#pragma unroll
		for (int q = 0; q < 8; q++) {
			uint4 w = src[q];
			a[4 * q] = w.x; a[4 * q + 1] = w.y; a[4 * q + 2] = w.z; a[4 * q + 3] = w.w;
		}
		if (!A.in_sliced) transpose32(a);
#pragma unroll
		for (int p = 0; p < 32; p++) tile[p * NU + uid] = a[p];
	}
	__syncthreads();

	const uint32_t n_steps = A.n_intra + A.n_inter;
	for (uint32_t step = 0; step < n_steps; step++) {
		// forward: inter-unit layers descending, then intra descending; inverse: the reverse order
		uint32_t layer = A.inverse ? step : (n_steps - 1 - step);
		if (layer >= A.n_intra) {
			uint32_t li = layer - A.n_intra + (L0 - A.n_intra);  // == layer - L0 when n_intra == L0
			li = layer - A.n_intra;                              // n_inter > 0 implies n_intra == L0
			for (uint32_t q = threadIdx.x; q < (NU >> 1); q += blockDim.x) {
				uint32_t kk = q & ((1u << li) - 1), jr = q >> li;
				uint32_t iu = (jr << (li + 1)) | kk, iv = iu | (1u << li);
				const uint32_t te = (1u << (Rt - 1 - li)) + jr;
				const uint32_t t = tw[te];
				const uint8_t *tb = twb + te * 32;
				uint32_t v[32], p[32];
#pragma unroll
				for (int k = 0; k < 32; k++) v[k] = tile[k * NU + iv];
				if (!A.inverse) {
					bs_mul_scalar(v, t, tb, p);
#pragma unroll
					for (int k = 0; k < 32; k++) {
						uint32_t u = tile[k * NU + iu] ^ p[k];
						tile[k * NU + iu] = u;
						tile[k * NU + iv] = v[k] ^ u;
					}
				} else {
#pragma unroll
					for (int k = 0; k < 32; k++) {
						v[k] ^= tile[k * NU + iu];
						tile[k * NU + iv] = v[k];
					}
					bs_mul_scalar(v, t, tb, p);
#pragma unroll
					for (int k = 0; k < 32; k++) tile[k * NU + iu] ^= p[k];
				}
			}
		} else {
			const uint32_t i = layer;
			const uint32_t S = 1u << (A.log_x + i);
			const uint32_t M = S == 1 ? 0x55555555u : S == 2 ? 0x33333333u : S == 4 ? 0x0F0F0F0Fu : S == 8 ? 0x00FF00FFu : 0x0000FFFFu;
			const uint32_t *tli = tl + i * 32;
			if (NU >= 2) {
				for (uint32_t q = threadIdx.x; q < (NU >> 1); q += blockDim.x) {
					uint32_t ia = 2 * q, ib = 2 * q + 1;
					uint32_t ta = thi[i * NU + ia], tb = thi[i * NU + ib];
					uint32_t v[32], t[32], p[32];
#pragma unroll
					for (int k = 0; k < 32; k++) {
						uint32_t wa = tile[k * NU + ia], wb = tile[k * NU + ib];
						if (A.inverse) {  // v += u first
							wa ^= (wa & M) << S;
							wb ^= (wb & M) << S;
							tile[k * NU + ia] = wa;
							tile[k * NU + ib] = wb;
						}
						v[k] = ((wa >> S) & M) | (wb & ~M);
						uint32_t ma = 0u - ((ta >> k) & 1u), mb = 0u - ((tb >> k) & 1u);
						t[k] = tli[k] ^ ((ma & M) | (mb & ~M));
					}
					BS<32>::mul(v, t, p);
#pragma unroll
					for (int k = 0; k < 32; k++) {
						uint32_t wa = tile[k * NU + ia] ^ (p[k] & M);
						uint32_t wb = tile[k * NU + ib] ^ ((p[k] & ~M) >> S);
						if (!A.inverse) {  // v += u after
							wa ^= (wa & M) << S;
							wb ^= (wb & M) << S;
						}
						tile[k * NU + ia] = wa;
						tile[k * NU + ib] = wb;
					}
				}
			} else if (threadIdx.x == 0) {  // a single unit in the tile: half of the lanes idle
				uint32_t ta = thi[i * NU];
				uint32_t v[32], t[32], p[32];
#pragma unroll
				for (int k = 0; k < 32; k++) {
					uint32_t wa = tile[k * NU];
					if (A.inverse) {
						wa ^= (wa & M) << S;
						tile[k * NU] = wa;
					}
					v[k] = (wa >> S) & M;
					t[k] = (tli[k] ^ (0u - ((ta >> k) & 1u))) & M;
				}
				BS<32>::mul(v, t, p);
#pragma unroll
				for (int k = 0; k < 32; k++) {
					uint32_t wa = tile[k * NU] ^ (p[k] & M);
					if (!A.inverse) wa ^= (wa & M) << S;
					tile[k * NU] = wa;
				}
			}
		}
		__syncthreads();
	}
	for (uint32_t uid = threadIdx.x; uid < NU; uid += blockDim.x) {
		uint4 *dst = reinterpret_cast<uint4 *>(base + (uint64_t)uid * 32);
		uint32_t a[32];
#pragma unroll
		for (int p = 0; p < 32; p++) a[p] = tile[p * NU + uid];
		if (!A.out_sliced) transpose32(a);
#pragma unroll
		for (int q = 0; q < 8; q++) dst[q] = make_uint4(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
	}
}

}  // namespace b200
