// roundevals_tc.cuh -- sumcheck round evaluations (inner products over GF(2^128)) on the 5th-gen tensor cores.
//
//   y = sum_i a_i * b_i  over GF(2^128)  is bilinear over GF(2):
//       sum_i a_i*b_i = sum_{p,q} parity(G[p][q]) * beta_p*beta_q ,   G[p][q] = sum_i bit_p(a_i) & bit_q(b_i)
//   G = A^T B is a dense {0,1} GEMM with M = N = 128 and K = number of hypercube points; only the
//   PARITY of every accumulator is needed, and only once per launch.  The kernel spreads the bits
//   of the operands over bytes in shared memory (MN-major, 128B swizzle), issues
//   tcgen05.mma.kind::i8 (UTCIMMA) with int32 accumulators in TMEM, and after the K loop reads the
//   accumulators back with tcgen05.ld, keeps bit 0, and XOR-combines the 128x128-bit matrices of all
//   CTAs in global memory.  A tiny second kernel folds G back into the field:
//       sum_p beta_p * (row p of G read as a field element)        (128 multiplications by basis elements)
//   mma.sync ... b1 is NOT usable for this on sm_100a: ptxas lowers it to IMMA.16832.U8 after an
//   in-register unpack, so the int8 path with our own unpack is the native one.
//
// Reference semantics: core/src/protocols/sumcheck/v3/bivariate_product.rs:303-408
//   y_1   = sum_c alpha^c sum_i hi_a*hi_b,   y_inf = sum_c alpha^c sum_i (lo_a+hi_a)*(lo_b+hi_b)
#pragma once
#include "field.cuh"
#include "linmap.cuh"

namespace b200 {
namespace tc {

constexpr uint32_t CHUNK = 64;                    // hypercube points per stage (= 2 MMA K-slices of 32)
constexpr uint32_t STAGE_BYTES = 4 * CHUNK * 128; // 4 operand tiles (a_hi, b_hi, a_inf, b_inf) of 64 x 128 bytes
constexpr uint32_t NSTAGE = 3;
constexpr uint32_t PRODUCERS = 256;               // one (point, operand) pair per producer thread and stage
constexpr uint32_t THREADS = PRODUCERS + 32;      // + one MMA-issuing warp (warp specialisation, no CTA-wide barriers)
constexpr uint32_t TMEM_COLS = 256;               // two 128x128 int32 accumulators
// instruction descriptor, kind::i8 (cute::UMMA::InstrDescriptor): D = S32 (2 << 4), A/B = UINT8,
// A and B MN-major (bits 15, 16), N = 128 (>> 3 at bit 17), M = 128 (>> 4 at bit 24)
constexpr uint32_t IDESC = (2u << 4) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, MN-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor, version 1):
// canonical layout ((16,8,m),(8,k)):((1,16,LBO),(128,SBO)) bytes -- every K index (hypercube point)
// owns one 128-byte row holding its 128 MN bytes, 16-byte chunk j of row i stored at chunk j ^ (i & 7);
// 8 rows = one 1 KiB swizzle atom, SBO = 1024 between K blocks.  (The un-swizzled MN-major layout
// works too but the tensor core's operand fetch then hits 8-way bank conflicts: 3x slower.)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
	uint64_t d = 0;
	d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
	d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 16;  // LBO (between MN blocks of 128; M = N = 128 has one block)
	d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;  // SBO (between K blocks of 8 rows)
	d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
	d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
	return d;
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	uint32_t done, addr = smem_u32(bar);
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
					 : "=r"(done)
					 : "r"(addr), "r"(parity)
					 : "memory");
	} while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
				 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
				 :
				 : "r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate), "r"(0u)
				 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// bits of one B128 -> 128 bytes in {0,1}.  Position pos = (w*8 + s)*4 + j holds bit rho(pos) = 32w + s + 8j
// (two ALU ops per output word; the permutation rho is undone by the combine kernel).
__device__ __forceinline__ uint32_t rho(uint32_t pos) { return 32u * (pos >> 5) + ((pos >> 2) & 7u) + 8u * (pos & 3u); }

// row i (128 B) of the operand tile; 16-byte chunk m_blk = pos / 16 = (w*8 + s) / 4 lands at chunk
// m_blk ^ (i & 7).  `row` points at the row, `sw16` = (i & 7) << 4 (both loop-invariant per thread).
// The byte at position pos = (w*8 + s)*4 + j is NOT normalised to {0,1}: it is bit rho(pos) of x left
// in place, i.e. 0 or 2^s (one LOP3 per output word, no shifts).  Every product contributing to
// D[pa][pb] then carries the same weight 2^(s(pa)+s(pb)), so the parity of the coincidence count is
// bit s(pa)+s(pb) of the (wrapping) int32 accumulator instead of bit 0.
__device__ __forceinline__ void unpack_store(uint8_t *row, uint32_t sw16, uint4 x) {
	const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
	for (uint32_t k = 0; k < 4; k++) {
		uint4 lo = make_uint4(w[k] & 0x01010101u, w[k] & 0x02020202u, w[k] & 0x04040404u, w[k] & 0x08080808u);
		uint4 hi = make_uint4(w[k] & 0x10101010u, w[k] & 0x20202020u, w[k] & 0x40404040u, w[k] & 0x80808080u);
		*reinterpret_cast<uint4 *>(row + (((2 * k) << 4) ^ sw16)) = lo;
		*reinterpret_cast<uint4 *>(row + (((2 * k + 1) << 4) ^ sw16)) = hi;
	}
}

// One inner-product job:  S = sum_{i < len} (a0[i] ^ a1[i]) * (b0[i] ^ b1[i])   (a1 / b1 may be null).
//   bivariate round evals: y_1 job = (a+half, -, b+half, -),  y_inf job = (a+half, a, b+half, b)
//   eq-ind round evals   : (E, -, val_{c,p}, -)
struct TcJob {
	const uint4 *a0, *a1, *b0, *b1;
};
struct TcArgs {
	const TcJob *jobs;  // 2 * gridDim.y jobs (padded with a copy of the last job when the count is odd)
	uint64_t len;       // points per job, a multiple of CHUNK
	uint32_t *gmat;     // [2 * gridDim.y][128][4] words, zero-initialised: XOR of the parity matrices of all CTAs
};

// grid = (ctas_per_job_pair, n_job_pairs), block = 288 (8 producer warps + 1 MMA warp), dyn smem = NSTAGE * STAGE_BYTES + 1024
__global__ void __launch_bounds__(THREADS) k_pair_tc(const TcArgs A) {
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzle atoms are 1 KiB aligned
	__shared__ __align__(8) uint64_t full[NSTAGE], empty[NSTAGE], done_bar;
	__shared__ uint32_t tmem_base_s;
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t c = blockIdx.y;  // job pair
	const uint64_t half = A.len;

	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	if (tid == 0) {
		for (uint32_t s = 0; s < NSTAGE; s++) {
			mbar_init(&full[s], PRODUCERS / 32);  // one arrival per producer warp
			mbar_init(&empty[s], 1);              // tcgen05.commit
		}
		mbar_init(&done_bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = tmem_base_s;

	const uint64_t n_chunks = (half + CHUNK - 1) / CHUNK;
	const uint32_t my_chunks = blockIdx.x < n_chunks ? (uint32_t)((n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
	if (warp < PRODUCERS / 32) {
		// ---- producers: thread -> (operand tile, point): tiles 0/1 = A/B of job 2c, tiles 2/3 = A/B of job 2c+1
		const uint32_t op = tid >> 6, pt = tid & 63;
		// operand of this thread: job 2c + (op >> 1), side A (op even) or B (op odd)
		const TcJob J = A.jobs[2 * c + (op >> 1)];
		const uint4 *q0 = (op & 1) ? J.b0 : J.a0, *q1 = (op & 1) ? J.b1 : J.a1;
		const bool need_lo = q1 != nullptr;
		if (!need_lo) q1 = q0;  // read the same line twice (L1 hit); selected away below
		// Loads are UNCONDITIONAL and land directly in their ring slot: a predicated load makes ptxas
		// load into temporaries and move them right away, which stalls every iteration for a full
		// memory latency.  len is a multiple of CHUNK (checked by the host); iterations past the end
		// re-read the last valid chunk (clamped offset).
		uint64_t off = (uint64_t)blockIdx.x * CHUNK + pt;
		const uint64_t off_last = (n_chunks - 1) * CHUNK + pt, step = (uint64_t)gridDim.x * CHUNK;
		auto load = [&](uint4 &xh, uint4 &xl) {
			const uint64_t o = off < off_last ? off : off_last;
			xh = __ldg(q0 + o);
			xl = __ldg(q1 + o);
			off += step;
		};
		// register ring of raw loads, one slot per pipeline stage (slot u <-> stage u): the HBM latency
		// spans several stages, and stage / ring indices stay compile-time constants.  (A ring twice as deep -- loads
		// six stages ahead -- measured the same 0.186 ms at m = 8, n = 20: the long_scoreboard stalls ncu shows are
		// producers with nothing else to do, not the limiter; that is the shared-memory port, which carries every
		// operand byte twice -- 32 KiB per stage written by the producers, 32 KiB read by the tensor core -- against
		// 256 tensor cycles per stage: 2 : 1, hence the ~45 % tensor-pipe ceiling of the 8x bit-to-byte expansion.)
		uint4 rh[NSTAGE], rl[NSTAGE];
#pragma unroll
		for (uint32_t u = 0; u < NSTAGE; u++) load(rh[u], rl[u]);
		uint8_t *row = smem + op * (CHUNK * 128) + pt * 128;
		const uint32_t sw16 = (pt & 7) << 4;
		uint32_t phase = 1;  // parity of the previous completion of empty[]; first round needs no wait
		for (uint32_t it0 = 0; it0 < my_chunks; it0 += NSTAGE, phase ^= 1) {
#pragma unroll
			for (uint32_t u = 0; u < NSTAGE; u++) {
				const uint32_t it = it0 + u;
				if (it < my_chunks) {
					const uint4 x = need_lo ? (rh[u] ^ rl[u]) : rh[u];
					load(rh[u], rl[u]);
					if (it0 > 0) mbar_wait(&empty[u], phase);  // MMAs that read this stage are done
					unpack_store(row + u * STAGE_BYTES, sw16, x);
					asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
					__syncwarp();
					if (lane == 0) mbar_arrive(&full[u]);
				}
			}
		}
	} else if (lane == 0) {
		// ---- MMA issuer: one thread
		uint32_t phase = 0;
		const uint64_t desc0 = make_desc(smem_u32(smem));
		for (uint32_t it0 = 0; it0 < my_chunks; it0 += NSTAGE, phase ^= 1) {
#pragma unroll
			for (uint32_t u = 0; u < NSTAGE; u++) {
				const uint32_t it = it0 + u;
				if (it < my_chunks) {
					mbar_wait(&full[u], phase);
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
					for (uint32_t ks = 0; ks < CHUNK / 32; ks++) {
						// descriptors differ only in the 16-byte-granular start address: base + constant
						const uint64_t d0 = desc0 + ((u * STAGE_BYTES + ks * 4 * 1024) >> 4);  // 4 K-blocks of 8 points per MMA
						umma_i8(tmem, d0, d0 + ((CHUNK * 128) >> 4), (it | ks) > 0);
						umma_i8(tmem + 128, d0 + ((2 * CHUNK * 128) >> 4), d0 + ((3 * CHUNK * 128) >> 4), (it | ks) > 0);
					}
					umma_commit(&empty[u]);
				}
			}
		}
		umma_commit(&done_bar);  // tracks completion of everything issued so far
	}
	const uint32_t iter = my_chunks;
	if (warp < PRODUCERS / 32) {
		mbar_wait(&done_bar, 0);
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	}

	if (iter > 0 && warp < PRODUCERS / 32) {
		// epilogue: warps 0-3 read accumulator 0, warps 4-7 accumulator 1; thread owns TMEM lane
		// m = 32*(warp%4) + lane = row pos_a
		const uint32_t acc = warp >> 2, wq = warp & 3;
		uint32_t *grow = A.gmat + ((size_t)c * 2 + acc) * 512 + (32 * wq + lane) * 4;
#pragma unroll
		for (uint32_t q = 0; q < 4; q++) {
			uint32_t r[32];
			const uint32_t taddr = tmem + ((32u * wq) << 16) + acc * 128 + q * 32;
			asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
						 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
						 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
						 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
						   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
						   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
						   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
						 : "r"(taddr)
						 : "memory");
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			// parity of D[m][32q + k] sits at bit s(m) + s(32q + k), s(pos) = (pos >> 2) & 7
			const uint32_t s_row = ((32 * wq + lane) >> 2) & 7;
			uint32_t wbits = 0;
#pragma unroll
			for (int k = 0; k < 32; k++) wbits = __funnelshift_r(wbits, r[k] >> (s_row + ((k >> 2) & 7)), 1);
			if (wbits) atomicXor(grow + q, wbits);
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

// G (position-permuted parity matrices) -> field.  One block per TARGET t = (job, slot, coef):
//   slots[slot_t] ^= coef_t * S_{job_t},   S_j = sum_m beta_{rho(m)} * (row m of G_j read as a field element)
// (a job may feed several targets: a shared inner product used by several compositions).
// grid = n_targets, block = 128, dyn smem = FIELD_TABLE_BYTES
struct TcTarget {
	uint32_t job, slot;
	uint4 coef;
};
__global__ void __launch_bounds__(128) k_pair_tc_combine(const uint8_t *__restrict__ g_tables, const uint32_t *__restrict__ gmat,
														   const TcTarget *__restrict__ targets, uint4 *__restrict__ slots) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	__shared__ uint4 red[32];
	const TcTarget tg = targets[blockIdx.x];
	const uint32_t j = tg.job, m = threadIdx.x;
	const uint32_t *row = gmat + (size_t)j * 512 + m * 4;
	// row value: bit rho(pb) set iff G'[m][pb]
	uint32_t v[4] = {0, 0, 0, 0};
#pragma unroll
	for (uint32_t q = 0; q < 4; q++) {
		uint32_t wbits = row[q];
		while (wbits) {
			uint32_t k = __ffs(wbits) - 1;
			wbits &= wbits - 1;
			uint32_t bit = rho(32 * q + k);
			v[bit >> 5] |= 1u << (bit & 31);
		}
	}
	uint4 term = basis_image(make_uint4(v[0], v[1], v[2], v[3]), rho(m));  // beta_{rho(m)} * row
	term = block_xor(term, red);
	if (threadIdx.x == 0 && !is_zero(term)) {
		uint4 cf = tg.coef;
		bool one = cf.x == 1 && (cf.y | cf.z | cf.w) == 0;
		atomic_xor_u4(slots + tg.slot, one ? term : f_mul128(T, term, cf));
	}
}

// G of one or more jobs -> 128 field elements: out[q] = the element whose bit p is (XOR over the jobs of) G[p][q].
// This is the OUTER-PRODUCT use of the bit-GEMM: with a = a B128 vector v and b = the 128-bit words w of a bit-packed B1
// matrix,  out[q] = sum_j v[j] * bit_q(w[j])  -- fold_left / evaluate_partial_high of a B1 multilinear by a large tensor
// query down to 7 variables (ring-switch partial evaluations, core/src/ring_switch/prove.rs:147-208).  block = 128.
// grid = number of outputs; block b combines the jobs b * n_jobs .. and writes out[128 b ..].
__global__ void __launch_bounds__(128) k_tc_outer_combine(const uint32_t *__restrict__ gmat, uint32_t n_jobs, uint4 *__restrict__ out) {
	__shared__ uint32_t o[128][4];
	gmat += (size_t)blockIdx.x * n_jobs * 512;
	out += (size_t)blockIdx.x * 128;
	const uint32_t m = threadIdx.x, p = rho(m);
	o[m][0] = o[m][1] = o[m][2] = o[m][3] = 0;
	__syncthreads();
	for (uint32_t j = 0; j < n_jobs; j++)
#pragma unroll
		for (uint32_t q = 0; q < 4; q++) {
			uint32_t wbits = gmat[(size_t)j * 512 + m * 4 + q];
			while (wbits) {
				const uint32_t k = __ffs(wbits) - 1;
				wbits &= wbits - 1;
				atomicXor(&o[rho(32 * q + k)][p >> 5], 1u << (p & 31));
			}
		}
	__syncthreads();
	out[m] = make_uint4(o[m][0], o[m][1], o[m][2], o[m][3]);
}

// The same jobs on rounds too small for the tensor cores (len < 4096 points): one CTA per target evaluates its job
// sum_i (a0 + a1)[i] * (b0 + b1)[i] with the per-lane multiply, scales it by the target's coefficient and XORs it
// into the slot -- ONE launch for all compositions of a small round.  grid = n_targets, block = 256, dyn smem = FIELD_TABLE_BYTES
__global__ void __launch_bounds__(256) k_jobs_small(const uint8_t *__restrict__ g_tables, const TcJob *__restrict__ jobs, const TcTarget *__restrict__ targets,
													uint64_t len, uint4 *__restrict__ slots) {
	extern __shared__ __align__(128) uint8_t smem[];
	FieldTables T = load_field_tables(smem, g_tables);
	__shared__ uint4 red[32];
	const TcTarget tg = targets[blockIdx.x];
	const TcJob J = jobs[tg.job];
	uint4 acc = u4_zero();
	for (uint64_t i = threadIdx.x; i < len; i += blockDim.x) {
		uint4 a = __ldg(J.a0 + i), b = __ldg(J.b0 + i);
		if (J.a1) a ^= __ldg(J.a1 + i);
		if (J.b1) b ^= __ldg(J.b1 + i);
		acc ^= f_mul128(T, a, b);
	}
	acc = block_xor(acc, red);
	if (threadIdx.x == 0 && !is_zero(acc)) {
		const uint4 cf = tg.coef;
		const bool one = cf.x == 1 && (cf.y | cf.z | cf.w) == 0;
		atomic_xor_u4(slots + tg.slot, one ? acc : f_mul128(T, acc, cf));
	}
}

}  // namespace tc
}  // namespace b200
