// groestl.cuh -- Groestl-256 on the device: leaf digests, pair compressions and whole binary Merkle trees over
// device-resident codewords (the RS codeword and the FRI round oracles never cross PCIe to be hashed).
//
// Reference: crates/hash/src/groestl/{digest.rs:60-90, compression.rs:22-36} (Groestl256, Groestl256ByteCompression),
// crates/core/src/merkle_tree/binary_merkle_tree.rs:27-211 (leaves = digests of `batch_size`-element chunks in their
// 16-byte little-endian serialisation, inner layers = pair compressions, flattened with the root last).
// Written from the Groestl specification (sections 3.2-3.4): the state is 8 columns of 64 bits (row 0 in the most
// significant byte); one round = AddRoundConstant, then for every output column the XOR of 8 table entries
//     T_r[ S-box input byte of row r taken from column (j + shift[r]) mod 8 ],   T_r[x] = rotr64(T_0[x], 8 r),
// T_0[x] = the MixBytes column circ(02,02,03,04,05,03,05,07) * S(x) (SubBytes, ShiftBytes and MixBytes fused).
// Shared-memory layout for conflict-free data-dependent gathers: an LDS.64 is served per half-warp (16 lanes x 8 B), so
// lane l only ever touches the bank pair l % 16 -- every table is stored as 16 REPLICAS, entry e of replica r at byte
// e * 128 + r * 8.  Four tables (rotations by 0, 8, 16, 24 bits = 128 KiB) are stored; T_{r+4} is T_r with its 32-bit
// halves swapped, which costs nothing.  One thread hashes one leaf / one pair.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace b200 {
namespace groestl {

constexpr uint32_t THREADS = 512;
constexpr uint32_t REPL = 16;
constexpr uint32_t TBL = 256 * REPL * 8;  // one rotated table with its 16 replicas: 32 KiB
constexpr uint32_t SMEM = 4 * TBL;        // 128 KiB

// tbl: this lane's replica, [8][256] uint2 {lo, hi}
template <bool Q>
__device__ __forceinline__ void permutation(uint32_t (&hi)[8], uint32_t (&lo)[8], const uint8_t *tbl) {
	constexpr int SH[8] = {Q ? 1 : 0, Q ? 3 : 1, Q ? 5 : 2, Q ? 7 : 3, Q ? 0 : 4, Q ? 2 : 5, Q ? 4 : 6, Q ? 6 : 7};
#pragma unroll 1
	for (uint32_t r = 0; r < 10; r++) {
#pragma unroll
		for (uint32_t c = 0; c < 8; c++) {
			if (Q) {
				hi[c] = ~hi[c];
				lo[c] = ~lo[c] ^ ((c << 4) ^ r);
			} else {
				hi[c] ^= ((c << 4) ^ r) << 24;
			}
		}
		uint32_t nh[8], nl[8];
#pragma unroll
		for (int j = 0; j < 8; j++) {
			uint32_t ah = 0, al = 0;
#pragma unroll
			for (int rr = 0; rr < 8; rr++) {
				const int src = (j + SH[rr]) & 7;
				const uint32_t w = rr < 4 ? hi[src] : lo[src];
				const int sft = 8 * (3 - (rr & 3));  // byte rr of the big-endian column
				const uint32_t off = sft >= 7 ? (w >> (sft - 7)) & 0x7F80u : (w << 7) & 0x7F80u;  // entry * 128
				const uint2 t = *reinterpret_cast<const uint2 *>(tbl + (rr & 3) * TBL + off);
				al ^= rr < 4 ? t.x : t.y;  // rotation by 32 more bits = swapped halves
				ah ^= rr < 4 ? t.y : t.x;
			}
			nh[j] = ah;
			nl[j] = al;
		}
#pragma unroll
		for (int j = 0; j < 8; j++) hi[j] = nh[j], lo[j] = nl[j];
	}
}

__device__ __forceinline__ uint32_t bswap(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// h ^= P(h ^ m) ^ Q(m); m given as 16 little-endian-loaded words of the 64-byte block
__device__ __forceinline__ void compress(uint32_t (&hh)[8], uint32_t (&hl)[8], const uint32_t (&m)[16], const uint8_t *tbl) {
	uint32_t ph[8], pl[8], qh[8], ql[8];
#pragma unroll
	for (int c = 0; c < 8; c++) {
		qh[c] = bswap(m[2 * c]);
		ql[c] = bswap(m[2 * c + 1]);
		ph[c] = hh[c] ^ qh[c];
		pl[c] = hl[c] ^ ql[c];
	}
	permutation<false>(ph, pl, tbl);
	permutation<true>(qh, ql, tbl);
#pragma unroll
	for (int c = 0; c < 8; c++) {
		hh[c] ^= ph[c] ^ qh[c];
		hl[c] ^= pl[c] ^ ql[c];
	}
}

// the same with Q(m) given: a block that consists of padding only is identical for every leaf, so is its Q
__device__ __forceinline__ void compress_q_known(uint32_t (&hh)[8], uint32_t (&hl)[8], const uint32_t (&m)[16], const uint32_t (&qh)[8],
												 const uint32_t (&ql)[8], const uint8_t *tbl) {
	uint32_t ph[8], pl[8];
#pragma unroll
	for (int c = 0; c < 8; c++) {
		ph[c] = hh[c] ^ bswap(m[2 * c]);
		pl[c] = hl[c] ^ bswap(m[2 * c + 1]);
	}
	permutation<false>(ph, pl, tbl);
#pragma unroll
	for (int c = 0; c < 8; c++) {
		hh[c] ^= ph[c] ^ qh[c];
		hl[c] ^= pl[c] ^ ql[c];
	}
}

// T_0 (256 x 8 bytes, {lo, hi}) from global memory -> 4 rotated tables x 16 replicas in shared memory;
// returns the lane's replica (byte offset lane % 16 * 8)
__device__ __forceinline__ const uint8_t *load_tables(uint8_t *smem, const uint2 *__restrict__ t0) {
	for (uint32_t e = threadIdx.x; e < 4 * 256 * REPL; e += blockDim.x) {
		const uint32_t rep = e & (REPL - 1), x = (e >> 4) & 255u, rr = e >> 12;
		const uint2 v = __ldg(t0 + x);
		const uint64_t w = ((uint64_t)v.y << 32) | v.x, rot = rr ? (w >> (8 * rr)) | (w << (64 - 8 * rr)) : w;
		*reinterpret_cast<uint2 *>(smem + rr * TBL + x * 128 + rep * 8) = make_uint2((uint32_t)rot, (uint32_t)(rot >> 32));
	}
	__syncthreads();
	return smem + (threadIdx.x & (REPL - 1)) * 8;
}

// digests[i] = Groestl256(data[i * leaf_bytes .. (i + 1) * leaf_bytes)); leaf_bytes a multiple of 16
__global__ void __launch_bounds__(THREADS) k_groestl_leaves(const uint2 *__restrict__ t0, const uint4 *__restrict__ data, uint64_t n_leaves,
															uint32_t leaf_bytes, uint4 *__restrict__ digests) {
	extern __shared__ __align__(16) uint8_t smem[];
	const uint8_t *tbl = load_tables(smem, t0);
	const uint32_t n_full = leaf_bytes / 64, rem = leaf_bytes % 64;
	const uint64_t n_blocks = (uint64_t)n_full + (rem <= 55 ? 1 : 2);
	// The last block is padding only when the leaf ends on a block boundary (0x80, zeros, block count) or when the padding
	// spills into a block of its own (zeros, block count): the same block for every leaf, so Q of it is computed once per
	// thread instead of once per leaf -- one of the 11 permutations of a 256-byte leaf (codeword cosets of 16 elements).
	const bool pad_only = rem == 0 || rem > 55;
	uint32_t mpad[16], cqh[8], cql[8];
#pragma unroll
	for (int q = 0; q < 14; q++) mpad[q] = 0;
	if (rem == 0) mpad[0] = 0x80u;
	mpad[14] = bswap((uint32_t)(n_blocks >> 32));
	mpad[15] = bswap((uint32_t)n_blocks);
#pragma unroll
	for (int c = 0; c < 8; c++) cqh[c] = bswap(mpad[2 * c]), cql[c] = bswap(mpad[2 * c + 1]);
	if (pad_only) permutation<true>(cqh, cql, tbl);
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_leaves; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint4 *src = data + i * (leaf_bytes / 16);
		uint32_t hh[8] = {0, 0, 0, 0, 0, 0, 0, 0}, hl[8] = {0, 0, 0, 0, 0, 0, 0, 0x100};  // IV: 256 as a big-endian u64 in the last column
		uint32_t m[16];
		for (uint32_t b = 0; b < n_full; b++) {
#pragma unroll
			for (int q = 0; q < 4; q++) {
				const uint4 v = __ldg(src + 4 * b + q);
				m[4 * q] = v.x, m[4 * q + 1] = v.y, m[4 * q + 2] = v.z, m[4 * q + 3] = v.w;
			}
			compress(hh, hl, m, tbl);
		}
		// padding: 0x80, zeros, block count as a big-endian u64 at the end of the last block
#pragma unroll
		for (int q = 0; q < 4; q++) {
			uint4 v = make_uint4(0, 0, 0, 0);
			if ((uint32_t)(16 * q) < rem) v = __ldg(src + 4 * n_full + q);
			m[4 * q] = v.x, m[4 * q + 1] = v.y, m[4 * q + 2] = v.z, m[4 * q + 3] = v.w;
		}
		{
			const uint32_t w = rem / 4;  // rem is a multiple of 16: the 0x80 byte starts word w
			if (w < 16) m[w] = 0x80u;
		}
		if (rem == 0) {
			compress_q_known(hh, hl, mpad, cqh, cql, tbl);
		} else if (rem <= 55) {
			m[14] = bswap((uint32_t)(n_blocks >> 32));
			m[15] = bswap((uint32_t)n_blocks);
			compress(hh, hl, m, tbl);
		} else {
			compress(hh, hl, m, tbl);
			compress_q_known(hh, hl, mpad, cqh, cql, tbl);
		}
		// output transformation: last 32 bytes of P(h) ^ h
		uint32_t ph[8], pl[8];
#pragma unroll
		for (int c = 0; c < 8; c++) ph[c] = hh[c], pl[c] = hl[c];
		permutation<false>(ph, pl, tbl);
		digests[2 * i] = make_uint4(bswap(ph[4] ^ hh[4]), bswap(pl[4] ^ hl[4]), bswap(ph[5] ^ hh[5]), bswap(pl[5] ^ hl[5]));
		digests[2 * i + 1] = make_uint4(bswap(ph[6] ^ hh[6]), bswap(pl[6] ^ hl[6]), bswap(ph[7] ^ hh[7]), bswap(pl[7] ^ hl[7]));
	}
}

// out[i] = Groestl256ByteCompression(in[2i], in[2i+1]) = last 32 bytes of P(x) ^ x, x = the 64 bytes of the pair
__global__ void __launch_bounds__(THREADS) k_groestl_compress_pairs(const uint2 *__restrict__ t0, const uint4 *__restrict__ in, uint64_t n_pairs,
																	uint4 *__restrict__ out) {
	extern __shared__ __align__(16) uint8_t smem[];
	const uint8_t *tbl = load_tables(smem, t0);
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t xh[8], xl[8], ph[8], pl[8];
#pragma unroll
		for (int q = 0; q < 4; q++) {
			const uint4 v = __ldg(in + 4 * i + q);
			xh[2 * q] = bswap(v.x), xl[2 * q] = bswap(v.y), xh[2 * q + 1] = bswap(v.z), xl[2 * q + 1] = bswap(v.w);
		}
#pragma unroll
		for (int c = 0; c < 8; c++) ph[c] = xh[c], pl[c] = xl[c];
		permutation<false>(ph, pl, tbl);
		out[2 * i] = make_uint4(bswap(ph[4] ^ xh[4]), bswap(pl[4] ^ xl[4]), bswap(ph[5] ^ xh[5]), bswap(pl[5] ^ xl[5]));
		out[2 * i + 1] = make_uint4(bswap(ph[6] ^ xh[6]), bswap(pl[6] ^ xl[6]), bswap(ph[7] ^ xh[7]), bswap(pl[7] ^ xl[7]));
	}
}

}  // namespace groestl
}  // namespace b200
