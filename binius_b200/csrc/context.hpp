// context.hpp -- the per-device context behind the C ABI (include/binius_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/binius_b200.h"

// one deferred extrapolate_line (see b200_extrapolate_line): consecutive folds with the same challenge
// on disjoint slices are batched into ONE multi-segment launch (the executor contract allows deferral
// as long as store-to-load order is kept, compute/src/layer.rs:90-99)
struct b200_pending_lerp {
	uint8_t *e0;
	const uint8_t *e1;
	uint64_t n;
};

// one deferred fold_right of the byte-LUT route (see fold_mat): consecutive calls with the SAME query (the projection of
// every witness column onto the rounds after the univariate skip, prove/zerocheck.rs:416-434) go out as one launch that
// builds the query's table once per CTA instead of once per CTA and column
struct b200_pending_fold_right {
	const uint8_t *mat;
	uint8_t *out;
};

// one recorded KernelExecutor op of an open kernel scope (b200_kernel_scope_begin .. _end)
struct b200_expr;
struct b200_trace_op {
	int kind = 0;  // 0 decl_value, 1 sum_composition_evals, 2 add
	uint32_t slot = 0;
	uint64_t init[2] = {0, 0};
	std::vector<const void *> inputs;
	uint64_t row_len = 0;
	const b200_expr *expr = nullptr;
	uint64_t coeff[2] = {0, 0};
	uint32_t log_len = 0;
	const void *a = nullptr, *b = nullptr;
	void *dst = nullptr;
};
struct b200_local_chunk {
	uint8_t *p;
	uint64_t bytes, used;
};

struct b200_ctx {
	// every entry point locks the context: `&self` calls from several host threads (rayon `join`/`map`,
	// compute/src/layer.rs:115-131) serialise on it; a kernel scope holds it from begin to end
	std::recursive_mutex mu;
	int device = 0;
	int n_sms = 148;
	cudaStream_t stream = nullptr;
	cudaStream_t own_stream = nullptr;
	// copy streams + events of the host-buffer pipeline (b200_extrapolate_line_host), created lazily
	cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
	cudaEvent_t ev_in[4] = {}, ev_k[4] = {}, ev_out[4] = {};
	// true while side-stream uploads may occupy the host-to-device copy engine: argument blocks are then fetched by a
	// one-CTA kernel from the (mapped) pinned mirror instead of a copy that would queue behind the upload -- the engine
	// does not arbitrate between streams, it drains the side stream's queue first (measured: tools/mb_overlap.cu)
	bool side_uploads = false;
	std::string err;
	uint64_t launches = 0;
	std::vector<b200_pending_lerp> pending;
	uint64_t pending_z[2] = {0, 0};
	std::vector<b200_pending_fold_right> pending_fr;  // never non-empty together with `pending`
	const void *pending_fr_vec = nullptr;
	uint32_t pending_fr_lvl = 0;
	uint64_t pending_fr_n_out = 0;

	// field tables (64 KiB B8 product table + 256 B times-X_2 table), device global memory
	uint8_t *d_tables = nullptr;
	uint64_t *d_groestl_t0 = nullptr;  // Groestl T_0 table (256 x 8 bytes), groestl.cuh
	// deferred scalar results (OpValue slots)
	uint4 *d_results = nullptr;
	uint32_t n_results = 0;
	uint64_t *h_results = nullptr;  // pinned mirror of the first slots (b200_results_fetch reads through it)
	// small device scratch for argument arrays (pointer lists etc.), ring-allocated
	uint8_t *d_args = nullptr;
	uint64_t args_off = 0;
	// pinned host mirror used for staging small argument arrays
	uint8_t *h_args = nullptr;
	// general device scratch (grown on demand)
	uint8_t *d_scratch = nullptr;
	uint64_t scratch_bytes = 0;
	// kernel scope (accumulate_kernels / map_kernels, layer.rs:134-245): ops are recorded and lowered at scope end
	bool tracing = false;
	std::vector<b200_trace_op> trace;
	std::vector<b200_local_chunk> local_pool;            // KernelMemMap::Local scratch, grown by chunks, never moved
	std::vector<std::pair<uint8_t *, uint64_t>> locals;  // (base, bytes) of the locals of the open scope
	// host-mapped mailbox of the persistent sumcheck tail (b200_sumcheck_tail_*), allocated on first use
	uint8_t *h_tail_mb = nullptr, *d_tail_mb = nullptr;
	size_t tail_mb_bytes = 0;
	uint8_t *d_expand_ws = nullptr;  // tensor expansion: in' and T of the outer-product plan (2 x 2^12 elements)
	uint8_t *d_tail_ws = nullptr;  // device workspace of the grid variant (accumulators, barrier, challenge relay)
	bool tail_active = false;
	std::vector<void *> deferred_free;  // device releases that arrived while a tail was running
	// kernel-selection switches for A/B measurements (b200_ctx_set_tuning); defaults = production paths
	int tune_ntt = 0;              // 0 look-up-table + bit-sliced low layers, 1 bit-sliced only, 2 scalar tables
	uint32_t tune_ntt_log_cc = 7;  // columns per work item of a look-up-table pass
	int tune_ntt_byte = 1;         // byte tables for the widest-shared layers of a look-up-table pass
	int tune_ntt_cw = 16;          // compute warps per CTA of a look-up-table pass (8: two CTAs per SM everywhere)
	int tune_fold = 2;             // 2 TMA-staged K64, 1 K64, 0 LUT128
	int tune_round_evals_tc = 1;   // 1 tensor-core plans, 0 per-lane kernels, 2 materialised values only
	int tune_uni_generic = 0;      // 1 forces the generic univariate-skip kernel
	int tune_expand_outer = 1;     // 0: tensor expansion by the doubling chain only (k_expand_small + k_expand_k64 rounds)
	int tune_uni_linear = 1;       // 0: linear monomials of the univariate-skip round stay in k_uni_b8 (A/B, tests)
	int tune_tail_grid = 1;        // 0: the persistent sumcheck kernel always runs on one CTA
	int tune_tail_resident = 1;    // 0: the small rounds of the persistent sumcheck kernel stay in device memory
	int tune_tail_trace = 0;       // 1: b200_sumcheck_tail_finish prints CTA 0's per-round time stamps (debugging aid)
};

namespace b200 {

constexpr uint32_t MAX_RESULTS = 1u << 16;
constexpr uint32_t H_RESULTS = 4096;  // slots mirrored in pinned host memory
constexpr uint64_t ARGS_BYTES = 4u << 20;

inline int32_t fail(b200_ctx *ctx, int32_t code, const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	if (ctx) ctx->err = buf;
	return code;
}

#define B200_CUDA(ctx, call)                                                                         \
	do {                                                                                             \
		cudaError_t e__ = (call);                                                                    \
		if (e__ != cudaSuccess)                                                                      \
			return b200::fail(ctx, B200_ERR_DEVICE, "%s failed: %s", #call, cudaGetErrorString(e__)); \
	} while (0)

#define B200_LAUNCH_CHECK(ctx)                                                                 \
	do {                                                                                       \
		(ctx)->launches++;                                                                     \
		cudaError_t e__ = cudaGetLastError();                                                  \
		if (e__ != cudaSuccess)                                                                \
			return b200::fail(ctx, B200_ERR_DEVICE, "kernel launch failed at %s:%d: %s", __FILE__, \
							  __LINE__, cudaGetErrorString(e__));                              \
	} while (0)

// copy a small host array into the device argument ring; returns device pointer (stream-ordered)
int32_t stage_args(b200_ctx *ctx, const void *host, uint64_t bytes, void **dev_out);
int32_t ensure_scratch(b200_ctx *ctx, uint64_t bytes);

inline bool is_pow2(uint64_t x) { return x && !(x & (x - 1)); }
inline uint32_t ilog2(uint64_t x) {
	uint32_t r = 0;
	while (x >>= 1) r++;
	return r;
}
inline uint4 to_u4(const uint64_t v[2]) {
	return make_uint4((uint32_t)v[0], (uint32_t)(v[0] >> 32), (uint32_t)v[1], (uint32_t)(v[1] >> 32));
}

}  // namespace b200
