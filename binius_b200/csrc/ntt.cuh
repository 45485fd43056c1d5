// ntt.cuh -- additive (LCH14) NTT kernels over the binary tower fields B8/B16/B32.
//
// Reference semantics: crates/ntt/src/tests/reference.rs:68-160 (scalar spec),
// crates/ntt/src/single_threaded.rs:134-362 (layer order, coset/twiddle indexing),
// crates/ntt/src/additive_ntt.rs:8-27 (batched layout  idx = x | y << log_x | z << (log_x+log_y)).
//
//   forward, layer i descending:  u += v * t ; v += u
//   inverse, layer i ascending :  v += u     ; u += v * t
//   t = twiddle_{row0+i}( coset << (log_y-1-i) | j ),  j = y >> (i+1)
//
// A "pass" executes R consecutive layers [i_lo, i_lo+R) on a tile held in shared memory, so a
// transform of n_layers layers costs ceil(n_layers / R) round trips over HBM instead of n_layers.
#pragma once
#include "field.cuh"

namespace b200 {

template <typename S> struct NttField;
template <> struct NttField<uint32_t> {
	static __device__ __forceinline__ uint32_t mul(const FieldTables &T, uint32_t a, uint32_t b) { return f_mul32(T, a, b); }
};
template <> struct NttField<uint16_t> {
	static __device__ __forceinline__ uint32_t mul(const FieldTables &T, uint32_t a, uint32_t b) { return f_mul16(T, a, b); }
};
template <> struct NttField<uint8_t> {
	static __device__ __forceinline__ uint32_t mul(const FieldTables &T, uint32_t a, uint32_t b) { return f_mul8(T, a, b); }
};

struct NttPassArgs {
	void *data;
	uint32_t log_x, log_y;   // log_x already includes the extension-degree shift
	uint32_t i_lo, R;        // tile spans layers [i_lo, i_lo + R) ...
	uint32_t R_exec;         // ... of which only [i_lo, i_lo + R_exec) are executed (R_exec <= R)
	uint32_t log_c;          // tile width along the inner (contiguous) axis
	uint32_t row0;           // s_evals row of layer 0 = d - (log_y + coset_bits)
	uint32_t d;              // log domain size
	uint64_t coset;
	int inverse;
	const uint32_t *s_evals; // device [32][32]
};

// dyn smem = FIELD_TABLE_BYTES + 4 * 2^R (twiddles) + sizeof(S) * 2^(R + log_c)
template <typename S>
__global__ void __launch_bounds__(256) k_ntt_pass(const uint8_t *__restrict__ g_tables, NttPassArgs A) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	uint32_t *tw = reinterpret_cast<uint32_t *>(smem + FIELD_TABLE_BYTES);
	S *tile = reinterpret_cast<S *>(smem + FIELD_TABLE_BYTES + (4u << A.R));

	const uint32_t R = A.R, i_hi = A.i_lo + R, log_c = A.log_c;
	const uint32_t log_inner = A.log_x + A.i_lo;
	const uint32_t n_inner_chunks_log = log_inner - log_c;
	uint64_t bid = blockIdx.x;
	uint64_t chunk = bid & ((1ull << n_inner_chunks_log) - 1);
	uint64_t outer = bid >> n_inner_chunks_log;  // y bits [i_hi, log_y)
	S *base = reinterpret_cast<S *>(A.data) + ((uint64_t)blockIdx.y << (A.log_x + A.log_y)) + (outer << (i_hi + A.log_x)) + (chunk << log_c);

	// twiddles: heap layout, layer li (i = i_lo + li) has 2^(R-1-li) entries at offset 2^(R-1-li)
	for (uint32_t e = threadIdx.x + 1; e < (1u << R); e += blockDim.x) {
		uint32_t lvl_bits = 31 - __clz(e);       // = R-1-li
		uint32_t li = R - 1 - lvl_bits;
		uint32_t jr = e - (1u << lvl_bits);
		uint32_t i = A.i_lo + li;
		uint64_t j = (outer << (i_hi - i - 1)) | jr;
		uint64_t idx = (A.coset << (A.log_y - 1 - i)) | j;
		uint32_t row = A.row0 + i;
		uint32_t nb = A.d - 1 - row;
		const uint32_t *srow = A.s_evals + row * 32;
		uint32_t t = 0;
		for (uint32_t b = 0; b < nb; b++)
			if ((idx >> b) & 1) t ^= srow[b];
		tw[e] = t;
	}
	const uint32_t n_tile = 1u << (R + log_c);
	const uint32_t cmask = (1u << log_c) - 1;
	for (uint32_t e = threadIdx.x; e < n_tile; e += blockDim.x) {
		uint32_t r = e >> log_c, c = e & cmask;
		tile[e] = base[((uint64_t)r << log_inner) + c];
	}
	__syncthreads();
	const uint32_t n_bf = n_tile >> 1;
	for (uint32_t step = 0; step < A.R_exec; step++) {
		uint32_t li = A.inverse ? step : (A.R_exec - 1 - step);
		for (uint32_t bfi = threadIdx.x; bfi < n_bf; bfi += blockDim.x) {
			uint32_t c = bfi & cmask, q = bfi >> log_c;
			uint32_t kk = q & ((1u << li) - 1), jr = q >> li;
			uint32_t r0 = (jr << (li + 1)) | kk, r1 = r0 | (1u << li);
			uint32_t t = tw[(1u << (R - 1 - li)) + jr];
			uint32_t u = tile[(r0 << log_c) + c], v = tile[(r1 << log_c) + c];
			if (!A.inverse) {
				u ^= NttField<S>::mul(T, t, v);  // the twiddle is shared by many lanes: first operand (field.cuh)
				v ^= u;
			} else {
				v ^= u;
				u ^= NttField<S>::mul(T, t, v);  // the twiddle is shared by many lanes: first operand (field.cuh)
			}
			tile[(r0 << log_c) + c] = (S)u;
			tile[(r1 << log_c) + c] = (S)v;
		}
		__syncthreads();
	}
	for (uint32_t e = threadIdx.x; e < n_tile; e += blockDim.x) {
		uint32_t r = e >> log_c, c = e & cmask;
		base[((uint64_t)r << log_inner) + c] = tile[e];
	}
}

}  // namespace b200
