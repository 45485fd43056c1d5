// capi.cu -- C ABI (include/binius_b200.h): validation + kernel launches.  No CPU fallback: every
// compute entry point launches sm_100a kernels on the context's stream or fails.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <mutex>

#include <map>
#include <set>
#include <tuple>

#include "context.hpp"
#include "eqind_plan.hpp"
#include "host_field.hpp"
#include "kernels.cuh"
#include "fold_tma.cuh"
#include "ntt.cuh"
#include "ntt_bs.cuh"
#include "ntt_lut.cuh"
#include "groestl.cuh"
#include "roundevals_tc.cuh"
#include "univariate.cuh"
#include "tail_grid.cuh"
#include "uni_split.hpp"

using namespace b200;

// device copies of several expressions' steps made in one allocation and one copy (upload_exprs): freed with the last user
struct StepSlab {
	void *base = nullptr;
	uint32_t refs = 0;
};
struct b200_expr {
	b200_ctx *ctx = nullptr;
	std::vector<b200_expr_step> steps;
	b200_expr_step *d_steps = nullptr;
	StepSlab *slab = nullptr;  // non-null: d_steps points into slab->base
	uint64_t uid = 0;          // unique per compiled expression (keys the cached monomial plan)
	uint32_t n_vars = 0;
	plan::Poly poly;       // monomial expansion (eqind_plan.hpp), valid when poly_ok
	bool poly_ok = false;  // degree <= 2 and few terms
};

// persistent sumcheck tail (k_sumcheck_tail): host-mapped mailboxes + progress
struct b200_tail {
	b200_ctx *ctx = nullptr;
	uint8_t *h_mb = nullptr;  // pinned, mapped
	uint8_t *d_mb = nullptr;  // device alias
	uint32_t n_vars = 0, n_vals = 0, round_out = 0, round_in = 0;
	size_t off_chal = 0, off_status = 0, off_trace = 0;
};

struct b200_ntt {
	b200_ctx *ctx = nullptr;
	uint32_t kt = 5, d = 0;
	std::vector<std::vector<uint64_t>> s_evals;  // rows (host), values < 2^(2^kt)
	uint32_t *d_s_evals = nullptr;               // device [32][32]
	uint32_t *d_basis = nullptr;                 // device [32][32][32]: s_evals[row][bit] * 2^k (B32 only; ntt_lut.cuh)
};

namespace b200 {

__global__ void k_fetch_args(const uint4 *__restrict__ src, uint4 *__restrict__ dst, uint32_t n) {
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

int32_t stage_args(b200_ctx *ctx, const void *host, uint64_t bytes, void **dev_out) {
	uint64_t need = (bytes + 255) & ~255ull;
	if (need > ARGS_BYTES) return fail(ctx, B200_ERR_ALLOC, "argument block of %llu bytes too large", (unsigned long long)bytes);
	if (ctx->args_off + need > ARGS_BYTES) {
		B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		ctx->args_off = 0;
	}
	memcpy(ctx->h_args + ctx->args_off, host, bytes);
	if (ctx->side_uploads) {
		// pinned host memory is device-accessible under unified addressing (cudaMallocHost)
		k_fetch_args<<<1, 256, 0, ctx->stream>>>(reinterpret_cast<const uint4 *>(ctx->h_args + ctx->args_off), reinterpret_cast<uint4 *>(ctx->d_args + ctx->args_off), (uint32_t)((bytes + 15) / 16));
		B200_CUDA(ctx, cudaGetLastError());
	} else {
		B200_CUDA(ctx, cudaMemcpyAsync(ctx->d_args + ctx->args_off, ctx->h_args + ctx->args_off, bytes, cudaMemcpyHostToDevice, ctx->stream));
	}
	*dev_out = ctx->d_args + ctx->args_off;
	ctx->args_off += need;
	return B200_OK;
}

int32_t ensure_scratch(b200_ctx *ctx, uint64_t bytes) {
	if (bytes <= ctx->scratch_bytes) return B200_OK;
	if (ctx->d_scratch) {
		B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_scratch);
		ctx->d_scratch = nullptr;
		ctx->scratch_bytes = 0;
	}
	// grow with 25% headroom in 32 MiB steps: successive rounds need slightly different sizes and a
	// reallocation costs a stream sync + cudaFree + cudaMalloc (~10 ms)
	uint64_t want = ((bytes + bytes / 4) + (32ull << 20) - 1) & ~((32ull << 20) - 1);
	if (cudaMalloc(&ctx->d_scratch, want) != cudaSuccess) {
		cudaGetLastError();
		want = bytes;
		if (cudaMalloc(&ctx->d_scratch, want) != cudaSuccess) {
			cudaGetLastError();
			return fail(ctx, B200_ERR_ALLOC, "out of device memory (scratch %llu bytes)", (unsigned long long)bytes);
		}
	}
	ctx->scratch_bytes = want;
	return B200_OK;
}

// several small host arrays -> ONE staged H2D copy
struct ArgPack {
	std::vector<uint8_t> buf;
	size_t add(const void *p, size_t bytes) {
		size_t off = (buf.size() + 15) & ~(size_t)15;
		buf.resize(off + bytes);
		if (bytes) memcpy(buf.data() + off, p, bytes);
		return off;
	}
	int32_t commit(b200_ctx *ctx, uint8_t **base) {
		void *d;
		int32_t rc = stage_args(ctx, buf.data(), buf.size(), &d);
		*base = (uint8_t *)d;
		return rc;
	}
};

static uint32_t grid_for(const b200_ctx *ctx, uint64_t n, uint32_t threads, uint32_t per_sm) {
	uint64_t blocks = (n + threads - 1) / threads;
	uint64_t cap = (uint64_t)ctx->n_sms * per_sm;
	return (uint32_t)std::max<uint64_t>(1, std::min(blocks, cap));
}

static int32_t new_slot(b200_ctx *ctx, uint32_t *slot) {
	if (ctx->n_results >= MAX_RESULTS) return fail(ctx, B200_ERR_ALLOC, "out of result slots");
	*slot = ctx->n_results++;
	return B200_OK;
}

template <typename K>
static int32_t set_smem(b200_ctx *ctx, K kernel, uint32_t bytes) {
	B200_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
	return B200_OK;
}

}  // namespace b200

// Inner-product jobs on the tensor cores: gmat scratch, k_pair_tc over job pairs, combine into result slots.
static int32_t launch_tc_pairs(b200_ctx *ctx, std::vector<tc::TcJob> &jobs, uint64_t len, const std::vector<tc::TcTarget> &targets,
							   uint64_t scratch_offset = 0, uint32_t **gmat_out = nullptr) {
	if (jobs.empty() || (targets.empty() && !gmat_out)) return B200_OK;
	if (jobs.size() & 1) jobs.push_back(jobs.back());  // pad to a pair; the duplicate's result is ignored
	const uint32_t n_pairs = (uint32_t)(jobs.size() / 2);
	if (n_pairs > 65535) return fail(ctx, B200_ERR_INPUT_VALIDATION, "too many inner-product jobs");
	const uint64_t gbytes = (uint64_t)jobs.size() * 512 * 4;
	int32_t rc = ensure_scratch(ctx, scratch_offset + gbytes);
	if (rc) return rc;
	uint32_t *gmat = (uint32_t *)(ctx->d_scratch + scratch_offset);
	B200_CUDA(ctx, cudaMemsetAsync(gmat, 0, gbytes, ctx->stream));
	ArgPack pack;
	size_t o_j = pack.add(jobs.data(), sizeof(tc::TcJob) * jobs.size()), o_t = pack.add(targets.data(), sizeof(tc::TcTarget) * targets.size());
	uint8_t *dbase;
	if ((rc = pack.commit(ctx, &dbase))) return rc;
	tc::TcArgs T;
	T.jobs = (const tc::TcJob *)(dbase + o_j);
	T.len = len;
	T.gmat = gmat;
	const uint64_t n_chunks = len / tc::CHUNK;
	// 2 CTAs per SM are resident (256 TMEM columns each).  Split every job pair over gx CTAs so that the
	// number of waves times the work per CTA is smallest: cost(gx) = ceil(gx * n_pairs / resident) *
	// (n_chunks / gx + fixed), `fixed` = prologue + epilogue of a CTA in units of one 64-point chunk.
	// (100 pairs on 296 slots: gx = 2 leaves a third of the slots idle; gx = 29 runs 9.8 -> 10 waves.)
	const uint32_t resident = 2 * ctx->n_sms;
	uint32_t gx = 1;
	{
		const double fixed = 24.0;
		double best = 1e300;
		const uint32_t gmax = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(n_chunks / 16, 1), std::max<uint64_t>(64, 4 * (uint64_t)resident / n_pairs));
		for (uint32_t g = 1; g <= gmax; g++) {
			const uint64_t waves = ((uint64_t)g * n_pairs + resident - 1) / resident;
			const double cost = (double)waves * ((double)n_chunks / g + fixed);
			if (cost < best * 0.999) {
				best = cost;
				gx = g;
			}
		}
	}
	tc::k_pair_tc<<<dim3(gx, n_pairs), tc::THREADS, tc::NSTAGE * tc::STAGE_BYTES + 1024, ctx->stream>>>(T);
	B200_LAUNCH_CHECK(ctx);
	if (gmat_out) {  // the caller combines the parity matrices itself
		*gmat_out = gmat;
		return B200_OK;
	}
	tc::k_pair_tc_combine<<<(uint32_t)targets.size(), 128, FIELD_TABLE_BYTES, ctx->stream>>>(ctx->d_tables, gmat, (const tc::TcTarget *)(dbase + o_t), ctx->d_results);
	B200_LAUNCH_CHECK(ctx);
	return B200_OK;
}

// Monomial plan for eq-ind round evaluations at the points 1 and infinity (eqind_plan.hpp): every
// (composition, point) sum is a coefficient-weighted combination of inner products shared between the
// compositions; a degree-2 monomial x*y costs one elementwise product w = E . P_x per covering variable.
struct MonoPlan {
	struct AJob {
		uint32_t code;
		int32_t scaled;  // -1: the eq-indicator itself, else index of the scaled vector
		int32_t y;       // -1: the all-ones vector (constant term), else the multilinear
	};
	struct AScale {
		uint32_t code, x;
	};
	std::vector<AJob> ajobs;
	std::vector<AScale> scales;
	std::vector<tc::TcTarget> targets;  // slots relative to the call's first slot
	bool need_ones = false;
};
static int32_t eq_ind_monomial_plan(b200_ctx *ctx, const b200_dev_ptr *mls, uint64_t half, b200_dev_ptr eq_ind, const b200_expr *const *comps,
									const b200_expr *const *leads, uint32_t n_comp, const uint32_t *codes, uint32_t n_points, uint32_t first_slot) {
	typedef MonoPlan::AJob AJob;
	typedef MonoPlan::AScale AScale;
	// The plan depends on the expressions and the point codes only, and a prover asks for the same one every round (the
	// greedy cover below costs ~0.3 ms of host time for the 75 chi constraints of keccak): the last plan is kept.
	static std::mutex plan_mu;
	static std::vector<std::pair<std::vector<uint64_t>, std::shared_ptr<const MonoPlan>>> plan_cache;  // most recent first, <= 4
	std::vector<uint64_t> key;
	key.reserve(2 * n_comp + n_points + 2);
	key.push_back(n_comp), key.push_back(n_points);
	for (uint32_t c = 0; c < n_comp; c++) key.push_back(comps[c]->uid), key.push_back(leads[c]->uid);
	for (uint32_t p = 0; p < n_points; p++) key.push_back(codes[p]);
	std::shared_ptr<const MonoPlan> plan;
	{
		std::lock_guard<std::mutex> g(plan_mu);
		for (auto &e : plan_cache)
			if (e.first == key) plan = e.second;
	}
	if (!plan) {
	MonoPlan P;
	std::vector<AJob> &ajobs = P.ajobs;
	std::vector<AScale> &scales = P.scales;
	std::vector<tc::TcTarget> &targets = P.targets;
	std::map<std::tuple<uint32_t, int32_t, int32_t>, uint32_t> job_ix;
	bool &need_ones = P.need_ones;
	for (uint32_t code = 1; code <= 2; code++) {
		std::vector<uint32_t> pts;
		for (uint32_t p = 0; p < n_points; p++)
			if (codes[p] == code) pts.push_back(p);
		if (pts.empty()) continue;
		auto poly = [&](uint32_t c) -> const plan::Poly & { return code == 1 ? comps[c]->poly : leads[c]->poly; };
		// greedy vertex cover of the degree-2 monomial graph; covering variables get scaled by E
		std::vector<std::pair<uint32_t, uint32_t>> left;
		uint32_t n_var = 0;
		for (uint32_t c = 0; c < n_comp; c++)
			for (auto &t : poly(c))
				if (t.first.size() == 2) {
					left.push_back({t.first[0], t.first[1]});
					n_var = std::max(n_var, std::max(t.first[0], t.first[1]) + 1);
				}
		std::sort(left.begin(), left.end());
		left.erase(std::unique(left.begin(), left.end()), left.end());
		std::map<uint32_t, int32_t> cover;
		std::vector<uint32_t> cnt(n_var, 0);
		for (auto &e : left) {
			cnt[e.first]++;
			if (e.second != e.first) cnt[e.second]++;
		}
		while (!left.empty()) {
			uint32_t best = 0;
			for (uint32_t v = 1; v < n_var; v++)
				if (cnt[v] > cnt[best]) best = v;  // the lowest index among the most frequent variables
			cover[best] = (int32_t)scales.size();
			scales.push_back(AScale{code, best});
			size_t keep = 0;
			for (auto &e : left) {
				if (e.first == best || e.second == best) {
					cnt[e.first]--;
					if (e.second != e.first) cnt[e.second]--;
				} else
					left[keep++] = e;
			}
			left.resize(keep);
		}
		auto job = [&](int32_t scaled, int32_t y) {
			auto key = std::make_tuple(code, scaled, y);
			auto it = job_ix.find(key);
			if (it != job_ix.end()) return it->second;
			uint32_t j = (uint32_t)ajobs.size();
			ajobs.push_back(AJob{code, scaled, y});
			job_ix[key] = j;
			return j;
		};
		for (uint32_t c = 0; c < n_comp; c++)
			for (auto &t : poly(c)) {
				uint32_t j;
				if (t.first.empty()) {
					need_ones = true;
					j = job(-1, -1);
				} else if (t.first.size() == 1) {
					j = job(-1, (int32_t)t.first[0]);
				} else {
					auto cx = cover.find(t.first[0]);
					j = cx != cover.end() ? job(cx->second, (int32_t)t.first[1]) : job(cover.at(t.first[1]), (int32_t)t.first[0]);
				}
				uint64_t w[2] = {(uint64_t)t.second, (uint64_t)(t.second >> 64)};
				for (uint32_t p : pts) targets.push_back(tc::TcTarget{j, c * n_points + p, to_u4(w)});
			}
	}
	plan = std::make_shared<const MonoPlan>(std::move(P));
	std::lock_guard<std::mutex> g(plan_mu);
	plan_cache.insert(plan_cache.begin(), {key, plan});  // (a first round asks for the point at infinity only: two plans per prover)
	if (plan_cache.size() > 4) plan_cache.pop_back();
	}
	const std::vector<AJob> &ajobs = plan->ajobs;
	const std::vector<AScale> &scales = plan->scales;
	std::vector<tc::TcTarget> targets = plan->targets;
	for (auto &t : targets) t.slot += first_slot;
	bool need_ones = plan->need_ones;
	if (targets.empty()) return B200_OK;  // all compositions vanish identically: slots stay zero
	// regular (unweighted) evaluator: E is the all-ones vector, so a "scaled" vector is the operand itself
	// (code 1) or hi + lo (code 2, formed by the job's second pointer) and no product kernel runs
	const bool weighted = eq_ind != nullptr;
	if (!weighted) need_ones = true;
	const uint64_t n_scale = weighted ? scales.size() : 0;
	const uint64_t off_ones = n_scale * half * 16, off_g = off_ones + (need_ones ? half * 16 : 0);
	int32_t rc = ensure_scratch(ctx, off_g + (uint64_t)(ajobs.size() + 1) * 512 * 4);
	if (rc) return rc;
	uint4 *base = (uint4 *)ctx->d_scratch, *ones = (uint4 *)(ctx->d_scratch + off_ones);
	if (n_scale) {
		std::vector<EqScaleOp> ops(n_scale);
		for (uint64_t s = 0; s < n_scale; s++) {
			const uint4 *x = (const uint4 *)mls[scales[s].x];
			ops[s] = EqScaleOp{x + half, scales[s].code == 2 ? x : nullptr, base + s * half};
		}
		void *d_ops;
		if ((rc = stage_args(ctx, ops.data(), sizeof(EqScaleOp) * n_scale, &d_ops))) return rc;
		uint32_t gx = grid_for(ctx, half, 256, 3);
		gx = std::max<uint32_t>(1, std::min<uint64_t>(gx, (uint64_t)ctx->n_sms * 6 / std::min<uint64_t>(n_scale, ctx->n_sms * 6) + 1));
		k_eq_scale<<<dim3(gx, (uint32_t)n_scale), 256, FIELD_TABLE_BYTES, ctx->stream>>>(ctx->d_tables, (const uint4 *)eq_ind, (const EqScaleOp *)d_ops, half);
		B200_LAUNCH_CHECK(ctx);
	}
	if (need_ones) {
		k_fill<<<grid_for(ctx, half, 256, 8), 256, 0, ctx->stream>>>(ones, half, make_uint4(1, 0, 0, 0));
		B200_LAUNCH_CHECK(ctx);
	}
	std::vector<tc::TcJob> jobs(ajobs.size());
	for (size_t j = 0; j < ajobs.size(); j++) {
		const AJob &a = ajobs[j];
		const uint4 *y = a.y < 0 ? nullptr : (const uint4 *)mls[a.y];
		const uint4 *a0, *a1 = nullptr;
		if (a.scaled < 0) {
			a0 = weighted ? (const uint4 *)eq_ind : ones;
		} else if (weighted) {
			a0 = base + (uint64_t)a.scaled * half;
		} else {
			const uint4 *x = (const uint4 *)mls[scales[a.scaled].x];
			a0 = x + half;
			a1 = a.code == 2 ? x : nullptr;
		}
		jobs[j] = tc::TcJob{a0, a1, y ? y + half : ones, (y && a.code == 2) ? y : nullptr};
	}
	return launch_tc_pairs(ctx, jobs, half, targets, off_g);
}

static int32_t flush_pending(b200_ctx *ctx);
// cudaFree synchronises with ALL running device work, i.e. it would block on a persistent sumcheck tail whose next step
// needs this very host thread: releases that arrive while a tail runs (typically a garbage-collected expression or NTT
// handle) are parked and performed when the tail finishes.
// (handles may outlive their context -- a garbage-collected wrapper released after b200_ctx_destroy --, so the
// context is only touched while it is in the registry of live contexts)
static std::mutex g_live_mu;
static std::set<b200_ctx *> g_live;
static void ctx_free(b200_ctx *ctx, void *p) {
	if (!p) return;
	{
		std::lock_guard<std::mutex> reg(g_live_mu);
		if (ctx && g_live.count(ctx)) {
			std::lock_guard<std::recursive_mutex> g(ctx->mu);
			if (ctx->tail_active) {
				ctx->deferred_free.push_back(p);
				return;
			}
		}
	}
	cudaFree(p);
}
// Every entry point takes the context lock (a context may be shared by several host threads: the trait methods
// take `&self`, compute/src/layer.rs:115-131) and makes the context's device current for the call.
struct CtxGuard {
	b200_ctx *c;
	int prev = -1;
	// recording = true: an entry point that only RECORDS while a kernel scope is open (decl_value / sum / add / local: ~600
	// calls per round of the PIOP sumcheck) -- it touches no device state then, so the device switch is skipped
	explicit CtxGuard(b200_ctx *ctx, bool recording = false) : c(ctx) {
		if (!c) return;
		c->mu.lock();
		if (recording && c->tracing && c->pending.empty() && c->pending_fr.empty()) return;  // (a queued fold would be launched by the call)
		if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
		if (prev != c->device) cudaSetDevice(c->device);
		else prev = -1;
	}
	~CtxGuard() {
		if (!c) return;
		if (prev >= 0) cudaSetDevice(prev);
		c->mu.unlock();
	}
	CtxGuard(const CtxGuard &) = delete;
};
#define B200_LOCK(ctx) CtxGuard guard__(ctx)
#define B200_LOCK_REC(ctx) CtxGuard guard__(ctx, true)
#define B200_FLUSH(ctx)                                  \
	do {                                                  \
		if ((ctx) && (ctx)->tail_active)                  \
			return b200::fail(ctx, B200_ERR_INPUT_VALIDATION, "a persistent sumcheck tail owns the stream: only b200_sumcheck_tail_* calls until it is finished"); \
		if ((ctx) && (!(ctx)->pending.empty() || !(ctx)->pending_fr.empty())) { \
			int32_t rc__ = flush_pending(ctx);            \
			if (rc__) return rc__;                        \
		}                                                 \
	} while (0)

// =================================================================================================
extern "C" {

int32_t b200_ctx_create(int32_t device, b200_ctx **out) {
	if (!out) return B200_ERR_INPUT_VALIDATION;
	*out = nullptr;
	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
		cudaGetLastError();
		return B200_ERR_DEVICE;
	}
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return B200_ERR_DEVICE;
	if (prop.major != 10) return B200_ERR_DEVICE;  // sm_100a only: there is no other code path
	std::unique_ptr<b200_ctx> ctx(new b200_ctx);
	ctx->device = device;
	ctx->n_sms = prop.multiProcessorCount;
	if (cudaSetDevice(device) != cudaSuccess) return B200_ERR_DEVICE;
	if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) return B200_ERR_DEVICE;
	ctx->stream = ctx->own_stream;
	bool ok = cudaMalloc(&ctx->d_tables, FIELD_TABLE_BYTES) == cudaSuccess &&
			  cudaMalloc(&ctx->d_results, sizeof(uint4) * MAX_RESULTS) == cudaSuccess &&
			  cudaMalloc(&ctx->d_args, ARGS_BYTES) == cudaSuccess && cudaMallocHost(&ctx->h_args, ARGS_BYTES) == cudaSuccess &&
			  cudaMallocHost(&ctx->h_results, 16 * H_RESULTS) == cudaSuccess;
	if (!ok) return B200_ERR_ALLOC;
	// 8-bit tables from the bit-level tower recursion (host_field.hpp)
	std::vector<uint8_t> tables(FIELD_TABLE_BYTES);
	for (int a = 0; a < 256; a++) {
		for (int b = a; b < 256; b++) {
			uint8_t p = (uint8_t)hostf::Tw<3>::mul((hostf::u128)a, (hostf::u128)b);
			tables[(a << 8) | b] = p;
			tables[(b << 8) | a] = p;
		}
		tables[65536 + a] = (uint8_t)hostf::Tw<3>::mul_alpha((hostf::u128)a);
	}
	if (cudaMemcpy(ctx->d_tables, tables.data(), FIELD_TABLE_BYTES, cudaMemcpyHostToDevice) != cudaSuccess) return B200_ERR_DEVICE;
	{
		// Groestl T_0: the MixBytes column circ(02,02,03,04,05,03,05,07) * S(x), row 0 in the most significant byte;
		// S = the AES S-box (inverse in GF(2^8) mod x^8+x^4+x^3+x+1, then the affine map)
		auto gmul = [](uint32_t a, uint32_t b) {
			uint32_t r = 0;
			for (; b; b >>= 1, a = ((a << 1) ^ ((a & 0x80) ? 0x11Bu : 0u)) & 0xFFu)
				if (b & 1) r ^= a;
			return r;
		};
		static const uint32_t MIX[8] = {2, 2, 3, 4, 5, 3, 5, 7};
		std::vector<uint64_t> t0(256);
		for (uint32_t x = 0; x < 256; x++) {
			uint32_t inv = 0;
			for (uint32_t y = 1; y < 256 && x; y++)
				if (gmul(x, y) == 1) inv = y;
			uint32_t sb = inv;
			for (int k = 1; k <= 4; k++) sb ^= ((inv << k) | (inv >> (8 - k))) & 0xFFu;
			sb ^= 0x63;
			uint64_t col = 0;
			for (uint32_t row = 0; row < 8; row++) col |= (uint64_t)gmul(MIX[(8 - row) & 7], sb) << (8 * (7 - row));
			t0[x] = col;
		}
		if (cudaMalloc(&ctx->d_groestl_t0, 2048) != cudaSuccess) return B200_ERR_ALLOC;
		if (cudaMemcpy(ctx->d_groestl_t0, t0.data(), 2048, cudaMemcpyHostToDevice) != cudaSuccess) return B200_ERR_DEVICE;
	}
	if (cudaMemset(ctx->d_results, 0, sizeof(uint4) * MAX_RESULTS) != cudaSuccess) return B200_ERR_DEVICE;
	b200_ctx *c = ctx.get();
	// opt in to large dynamic shared memory once
	int32_t rc = B200_OK;
#define SET(k, bytes) if (rc == B200_OK) rc = set_smem(c, k, bytes)
	// (function attributes are per device: set them for every context, not once per process)
	SET((k_lerp_tma<false, false>), FT_SMEM);
	SET((k_lerp_tma<true, false>), FT_SMEM);
	SET((k_lerp_tma<false, true>), FT_SMEM + FT_MAX_SEGS * sizeof(LerpSeg));
	SET((k_lerp_tma<true, true>), FT_SMEM + FT_MAX_SEGS * sizeof(LerpSeg));
	SET((k_lerp_lut<512, 2, 2, true>), LUT_BYTES + 2048);
	SET((k_lerp_lut<512, 2, 2, false>), LUT_BYTES + 2048);
	SET((k_lerp_pairs_lut<512, 2, 2, true>), LUT_BYTES + 2048);
	SET(k_expand_k64<1>, 1 * LUT_BYTES + 6144);
	SET(k_expand_k64<2>, 2 * LUT_BYTES + 6144);
	SET(k_expand_k64<3>, 3 * LUT_BYTES + 6144);
	SET(k_expand_small, EXP_SMALL_LOG * (NLUT_BYTES + 2048) + (16u << EXP_SMALL_LOG));
	SET(k_expand_pair, ((FIELD_TABLE_BYTES + 127) & ~127u) + 2 * (16u << EXP_SMALL_LOG));
	SET(k_expand_outer, LUT_BYTES + 2048 + (16u << EXP_SMALL_LOG));
	SET(k_inner_product, FIELD_TABLE_BYTES);
	SET(k_fold_mat<false>, FIELD_TABLE_BYTES);
	SET(k_fold_right_lut, LUT_BYTES + 2048);
	SET(k_fold_right_lut_multi, LUT_BYTES + 2048);
	SET(k_fold_left_b1_lut, LUT_BYTES + 2048);
	SET(k_linear_map, LUT_BYTES + 2048);
	SET(k_fold_mat<true>, FIELD_TABLE_BYTES);
	SET(k_compute_composite, FIELD_TABLE_BYTES);
	SET(k_sum_composition, FIELD_TABLE_BYTES);
	SET(k_pairwise_product, FIELD_TABLE_BYTES);
	SET(k_bivariate_round_evals, FIELD_TABLE_BYTES);
	SET(k_eq_ind_round_evals, FIELD_TABLE_BYTES);
	SET(k_eq_ind_vals, FIELD_TABLE_BYTES);
	SET(k_eq_scale, FIELD_TABLE_BYTES);
	SET(k_sumcheck_tail_grid, TG_SMEM_MAX);
	SET(k_fri_fold, FIELD_TABLE_BYTES);
	SET(k_fri_lerp_k64<1>, 1 * LUT_BYTES + 6144 + NLUT_BYTES);
	SET(k_fri_lerp_k64<2>, 2 * LUT_BYTES + 6144 + NLUT_BYTES);
	SET(k_fri_lerp_k64<3>, 3 * LUT_BYTES + 6144 + NLUT_BYTES);
	SET(k_fri_lerp_k64<4>, 3 * LUT_BYTES + 6144 + NLUT_BYTES);
	SET(k_fri_fold_lut<1>, FIELD_TABLE_BYTES + 1 * NLUT_BYTES);
	SET(k_fri_fold_lut<2>, FIELD_TABLE_BYTES + 2 * NLUT_BYTES);
	SET(k_fri_fold_lut<3>, FIELD_TABLE_BYTES + 3 * NLUT_BYTES);
	SET(k_fri_fold_lut<4>, FIELD_TABLE_BYTES + 4 * NLUT_BYTES);
	SET(k_fri_fold_lut<5>, FIELD_TABLE_BYTES + 5 * NLUT_BYTES);
	SET(k_ntt_pass<uint32_t>, FIELD_TABLE_BYTES + 4 * 8192 + 4 * 8192);
	SET(k_ntt_pass<uint16_t>, FIELD_TABLE_BYTES + 4 * 8192 + 2 * 8192);
	SET(k_ntt_pass<uint8_t>, FIELD_TABLE_BYTES + 4 * 8192 + 1 * 8192);
	SET(k_ntt_bs_pass, 36 * 1024 + 16 + 128 * 1024);
	SET(k_ntt_bs_low, 184 * 1024 + 640 + 16);
	SET((nttl::k_ntt_lut<0, false, false>), 227u * 1024u);
	SET((nttl::k_ntt_lut<0, true, false>), 227u * 1024u);
	SET((nttl::k_ntt_lut<1, false, false>), 227u * 1024u);
	SET((nttl::k_ntt_lut<1, true, false>), 227u * 1024u);
	SET((nttl::k_ntt_lut<2, false, false>), 227u * 1024u);
	SET((nttl::k_ntt_lut<2, false, true>), 227u * 1024u);
	SET((nttl::k_ntt_lut<2, true, false>), 227u * 1024u);
	SET((nttl::k_ntt_lut<2, true, true>), 227u * 1024u);
	SET((nttl::k_ntt_lut<3, false, false>), 227u * 1024u);
	SET((nttl::k_ntt_lut<3, false, true>), 227u * 1024u);
	SET((nttl::k_ntt_lut<3, true, false>), 227u * 1024u);
	SET((nttl::k_ntt_lut<3, true, true>), 227u * 1024u);
	SET(tc::k_pair_tc, tc::NSTAGE * tc::STAGE_BYTES + 1024);
	SET(tc::k_pair_tc_combine, FIELD_TABLE_BYTES);
	SET(tc::k_jobs_small, FIELD_TABLE_BYTES);
	SET(groestl::k_groestl_leaves, groestl::SMEM);
	SET(groestl::k_groestl_compress_pairs, groestl::SMEM);
#undef SET
	if (rc != B200_OK) return rc;
	*out = ctx.release();
	{
		std::lock_guard<std::mutex> reg(g_live_mu);
		g_live.insert(*out);
	}
	return B200_OK;
}

void b200_ctx_destroy(b200_ctx *ctx) {
	if (!ctx) return;
	{
		std::lock_guard<std::mutex> reg(g_live_mu);
		g_live.erase(ctx);
	}
	cudaSetDevice(ctx->device);
	if (!ctx->pending.empty() || !ctx->pending_fr.empty()) flush_pending(ctx);
	cudaStreamSynchronize(ctx->stream);
	cudaFree(ctx->d_tables);
	if (ctx->d_groestl_t0) cudaFree(ctx->d_groestl_t0);
	cudaFree(ctx->d_results);
	cudaFree(ctx->d_args);
	cudaFreeHost(ctx->h_args);
	cudaFreeHost(ctx->h_results);
	if (ctx->d_scratch) cudaFree(ctx->d_scratch);
	for (auto &c : ctx->local_pool) cudaFree(c.p);
	if (ctx->h_tail_mb) cudaFreeHost(ctx->h_tail_mb);
	if (ctx->d_tail_ws) cudaFree(ctx->d_tail_ws);
	if (ctx->d_expand_ws) cudaFree(ctx->d_expand_ws);
	if (ctx->s_h2d) {
		cudaStreamDestroy(ctx->s_h2d);
		cudaStreamDestroy(ctx->s_d2h);
		for (int i = 0; i < 4; i++) {
			cudaEventDestroy(ctx->ev_in[i]);
			cudaEventDestroy(ctx->ev_k[i]);
			cudaEventDestroy(ctx->ev_out[i]);
		}
	}
	cudaStreamDestroy(ctx->own_stream);
	delete ctx;
}

const char *b200_last_error(b200_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void *b200_ctx_stream(b200_ctx *ctx) {
	// a caller that orders its own work on the stream must see every queued fold launched
	if (!ctx) return nullptr;
	B200_LOCK(ctx);
	if (!ctx->pending.empty() || !ctx->pending_fr.empty()) flush_pending(ctx);
	return (void *)ctx->stream;
}
int32_t b200_flush(b200_ctx *ctx) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	return B200_OK;
}
uint64_t b200_ctx_launch_count(b200_ctx *ctx) { return ctx ? ctx->launches : 0; }
int32_t b200_ctx_set_stream(b200_ctx *ctx, void *s) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
	return B200_OK;
}
int32_t b200_ctx_set_tuning(b200_ctx *ctx, const char *key, int32_t value) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx || !key) return B200_ERR_INPUT_VALIDATION;
	if (!strcmp(key, "ntt")) ctx->tune_ntt = value;
	else if (!strcmp(key, "ntt_log_cc")) ctx->tune_ntt_log_cc = (uint32_t)std::min(7, std::max(5, value));
	else if (!strcmp(key, "ntt_cw")) ctx->tune_ntt_cw = value;
	else if (!strcmp(key, "ntt_byte")) ctx->tune_ntt_byte = value;
	else if (!strcmp(key, "fold")) ctx->tune_fold = value;
	else if (!strcmp(key, "round_evals_tc")) ctx->tune_round_evals_tc = value;
	else if (!strcmp(key, "uni_generic")) ctx->tune_uni_generic = value;
	else if (!strcmp(key, "uni_linear")) ctx->tune_uni_linear = value;
	else if (!strcmp(key, "expand_outer")) ctx->tune_expand_outer = value;
	else if (!strcmp(key, "tail_grid")) ctx->tune_tail_grid = value;
	else if (!strcmp(key, "tail_resident")) ctx->tune_tail_resident = value;
	else if (!strcmp(key, "tail_trace")) ctx->tune_tail_trace = value;
	else return fail(ctx, B200_ERR_INPUT_VALIDATION, "unknown tuning key %s", key);
	return B200_OK;
}
void b200_host_mul128(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]) {
	const hostf::u128 p = hostf::mul128(hostf::from_words(a), hostf::from_words(b));
	out[0] = (uint64_t)p;
	out[1] = (uint64_t)(p >> 64);
}
int32_t b200_event_create(b200_ctx *ctx, void **out) {
	B200_LOCK(ctx);
	if (!ctx || !out) return B200_ERR_INPUT_VALIDATION;
	cudaEvent_t e;
	B200_CUDA(ctx, cudaEventCreate(&e));
	*out = e;
	return B200_OK;
}
int32_t b200_event_record(b200_ctx *ctx, void *e) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	B200_CUDA(ctx, cudaEventRecord((cudaEvent_t)e, ctx->stream));
	return B200_OK;
}
int32_t b200_event_elapsed_ms(b200_ctx *ctx, void *a, void *b, float *ms) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx || !ms) return B200_ERR_INPUT_VALIDATION;
	B200_CUDA(ctx, cudaEventSynchronize((cudaEvent_t)b));
	B200_CUDA(ctx, cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
	return B200_OK;
}
void b200_event_destroy(void *e) {
	if (e) cudaEventDestroy((cudaEvent_t)e);
}

int32_t b200_dev_alloc(b200_ctx *ctx, uint64_t n_elems, b200_dev_ptr *out) {
	B200_LOCK(ctx);
	if (!ctx || !out) return B200_ERR_INPUT_VALIDATION;
	void *p = nullptr;
	if (cudaMalloc(&p, std::max<uint64_t>(n_elems, 1) * 16) != cudaSuccess) {
		cudaGetLastError();
		return fail(ctx, B200_ERR_ALLOC, "out of device memory allocating %llu elements", (unsigned long long)n_elems);
	}
	*out = p;
	return B200_OK;
}
int32_t b200_dev_free(b200_ctx *ctx, b200_dev_ptr p) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	B200_CUDA(ctx, cudaFree(p));
	return B200_OK;
}
int32_t b200_host_alloc(b200_ctx *ctx, uint64_t n_bytes, void **out) {
	B200_LOCK(ctx);
	if (!ctx || !out) return B200_ERR_INPUT_VALIDATION;
	if (cudaMallocHost(out, std::max<uint64_t>(n_bytes, 1)) != cudaSuccess) {
		cudaGetLastError();
		return fail(ctx, B200_ERR_ALLOC, "out of pinned host memory");
	}
	return B200_OK;
}
int32_t b200_host_free(b200_ctx *ctx, void *p) {
	B200_LOCK(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	B200_CUDA(ctx, cudaFreeHost(p));
	return B200_OK;
}

int32_t b200_copy_h2d(b200_ctx *ctx, const void *src, b200_dev_ptr dst, uint64_t n) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (n == 0) return B200_OK;
	B200_CUDA(ctx, cudaMemcpyAsync(dst, src, n * 16, cudaMemcpyHostToDevice, ctx->stream));
	// pageable sources are staged synchronously by the runtime; pinned ones stay async
	return B200_OK;
}
// Host-to-device copy on the context's SIDE stream: it overlaps with everything already (and subsequently) issued on the
// main stream until b200_side_join makes the main stream wait for it.  For streaming a witness in while earlier chunks
// are being processed (the caller keeps the destination untouched by main-stream work until the join).
static int32_t ensure_side_streams(b200_ctx *ctx) {
	if (ctx->s_h2d) return B200_OK;
	B200_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
	B200_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
	for (uint32_t i = 0; i < 4; i++) {
		B200_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
		B200_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming));
		B200_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
	}
	return B200_OK;
}
int32_t b200_copy_h2d_side(b200_ctx *ctx, const void *src, b200_dev_ptr dst, uint64_t n) {
	B200_LOCK(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (n == 0) return B200_OK;
	int32_t rc = ensure_side_streams(ctx);
	if (rc) return rc;
	B200_CUDA(ctx, cudaMemcpyAsync(dst, src, n * 16, cudaMemcpyHostToDevice, ctx->s_h2d));
	return B200_OK;
}
int32_t b200_side_join(b200_ctx *ctx) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (!ctx->s_h2d) return B200_OK;
	B200_CUDA(ctx, cudaEventRecord(ctx->ev_in[0], ctx->s_h2d));
	B200_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[0], 0));
	return B200_OK;
}
int32_t b200_copy_d2h(b200_ctx *ctx, b200_dev_ptr src, void *dst, uint64_t n) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (n) B200_CUDA(ctx, cudaMemcpyAsync(dst, src, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
	B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return B200_OK;
}
int32_t b200_copy_d2d(b200_ctx *ctx, b200_dev_ptr src, b200_dev_ptr dst, uint64_t n) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (n) B200_CUDA(ctx, cudaMemcpyAsync(dst, src, n * 16, cudaMemcpyDeviceToDevice, ctx->stream));
	return B200_OK;
}
int32_t b200_fill(b200_ctx *ctx, b200_dev_ptr dst, uint64_t n, const uint64_t value[2]) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (n == 0) return B200_OK;
	k_fill<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>((uint4 *)dst, n, to_u4(value));
	B200_LAUNCH_CHECK(ctx);
	return B200_OK;
}
int32_t b200_sync(b200_ctx *ctx) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return B200_OK;
}

int32_t b200_results_reset(b200_ctx *ctx) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (ctx->n_results) B200_CUDA(ctx, cudaMemsetAsync(ctx->d_results, 0, sizeof(uint4) * ctx->n_results, ctx->stream));
	ctx->n_results = 0;
	return B200_OK;
}
int32_t b200_results_fetch(b200_ctx *ctx, const uint32_t *slots, uint32_t n, uint64_t *host_out) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (n == 0) {
		B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		return B200_OK;
	}
	uint32_t hi = 0;
	for (uint32_t i = 0; i < n; i++) {
		if (slots[i] >= ctx->n_results) return fail(ctx, B200_ERR_INPUT_VALIDATION, "result slot %u out of range", slots[i]);
		hi = std::max(hi, slots[i]);
	}
	// the round values of a sumcheck round are a handful of slots: read them through the pinned mirror
	// (a pageable destination costs an extra staging copy inside the driver, ~10 us per round)
	std::vector<uint64_t> big;
	uint64_t *tmp = ctx->h_results;
	if (hi + 1 > H_RESULTS) {
		big.resize(2 * (size_t)(hi + 1));
		tmp = big.data();
	}
	B200_CUDA(ctx, cudaMemcpyAsync(tmp, ctx->d_results, 16 * (size_t)(hi + 1), cudaMemcpyDeviceToHost, ctx->stream));
	B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	for (uint32_t i = 0; i < n; i++) {
		host_out[2 * i] = tmp[2 * slots[i]];
		host_out[2 * i + 1] = tmp[2 * slots[i] + 1];
	}
	return B200_OK;
}

// -------------------------------------------------------------------------------------------------
}  // extern "C"
template <uint32_t THREADS, uint32_t UNR, int MINB, bool PAIRS = false, bool K64 = true>
static int32_t launch_lerp_variant(b200_ctx *ctx, const std::vector<LerpSeg> &live, const uint64_t z[2]) {
	constexpr uint32_t TILE = THREADS * UNR;
	auto kernel = PAIRS ? k_lerp_pairs_lut<THREADS, UNR, MINB, K64> : k_lerp_lut<THREADS, UNR, MINB, K64>;  // smem opt-in: b200_ctx_create
	// segments travel by value in the kernel parameters: no staging copy, one launch per <= 48 segments
	for (size_t s0 = 0; s0 < live.size(); s0 += LERP_MAX_SEGS) {
		LerpArgs A;
		A.segs_dev = nullptr;
		A.n_segs = (uint32_t)std::min<size_t>(LERP_MAX_SEGS, live.size() - s0);
		uint64_t tiles = 0;
		for (uint32_t i = 0; i < A.n_segs; i++) {
			A.segs[i] = live[s0 + i];
			A.segs[i].tile_start = tiles;
			tiles += (A.segs[i].upper + TILE - 1) / TILE;
		}
		A.n_tiles = tiles;
		A.z = to_u4(z);
		uint32_t grid = (uint32_t)std::min<uint64_t>(tiles, (uint64_t)ctx->n_sms * MINB);
		kernel<<<grid, THREADS, LUT_BYTES + 2048, ctx->stream>>>(A);
		B200_LAUNCH_CHECK(ctx);
	}
	return B200_OK;
}

// TMA-staged persistent kernel (fold_tma.cuh): tiles of FT_TILE outputs, one CTA per SM
template <bool PAIRS>
static int32_t launch_lerp_tma(b200_ctx *ctx, const std::vector<LerpSeg> &live, const uint64_t z[2]) {
	// one launch per FT_MAX_SEGS segments: up to LERP_MAX_SEGS by value, longer lists staged in device
	// memory (the kernel copies them to shared memory)
	for (size_t s0 = 0; s0 < live.size(); s0 += FT_MAX_SEGS) {
		const size_t cnt = std::min<size_t>(FT_MAX_SEGS, live.size() - s0);
		LerpArgs A;
		std::vector<LerpSeg> all;
		const bool by_value = cnt <= LERP_MAX_SEGS;
		if (!by_value) all.resize(cnt);
		LerpSeg *dst = by_value ? A.segs : all.data();
		uint64_t tiles = 0;
		for (size_t i = 0; i < cnt; i++) {
			dst[i] = live[s0 + i];
			dst[i].tile_start = tiles;
			tiles += (live[s0 + i].upper + FT_TILE - 1) / FT_TILE;
		}
		if (tiles >> 32) return fail(ctx, B200_ERR_INPUT_VALIDATION, "fold: too many tiles in one call");
		A.segs_dev = nullptr;
		if (!by_value) {
			void *d;
			int32_t rc = stage_args(ctx, all.data(), sizeof(LerpSeg) * cnt, &d);
			if (rc) return rc;
			A.segs_dev = (const LerpSeg *)d;
		}
		A.n_segs = (uint32_t)cnt;
		A.n_tiles = tiles;
		A.z = to_u4(z);
		uint32_t grid = (uint32_t)std::min<uint64_t>((tiles + FT_WARPS - 1) / FT_WARPS, (uint64_t)ctx->n_sms);
		const uint32_t smem = FT_SMEM + (by_value ? 0 : (uint32_t)(sizeof(LerpSeg) * cnt));
		if (by_value) k_lerp_tma<PAIRS, false><<<grid, FT_THREADS, smem, ctx->stream>>>(A);
		else k_lerp_tma<PAIRS, true><<<grid, FT_THREADS, smem, ctx->stream>>>(A);
		B200_LAUNCH_CHECK(ctx);
	}
	return B200_OK;
}
extern "C" {
static int32_t launch_lerp(b200_ctx *ctx, std::vector<LerpSeg> &segs, const uint64_t z[2]) {
	std::vector<LerpSeg> live;
	for (auto &s : segs)
		if (s.upper) {
			live.push_back(s);
		}
	if (live.empty()) return B200_OK;
	// 512 threads x 2 elements in flight, 2 CTAs per SM (64 registers): best of the measured variants.
	// B200_FOLD_ENGINE=lut128 selects the 16 x LDS.128 engine instead of the Karatsuba-64 one (A/B runs).
	const int engine = ctx->tune_fold;
	if (engine == 0) return launch_lerp_variant<512, 2, 2, false, false>(ctx, live, z);
	if (engine == 1) return launch_lerp_variant<512, 2, 2, false, true>(ctx, live, z);
	for (auto &sg : live)
		if (sg.upper >> 32) return launch_lerp_variant<512, 2, 2, false, true>(ctx, live, z);  // 32-bit tile arithmetic in the TMA kernel
	return launch_lerp_tma<false>(ctx, live, z);
}

int32_t b200_extrapolate_line(b200_ctx *ctx, b200_dev_ptr e0, uint64_t n0, b200_dev_ptr e1, uint64_t n1, const uint64_t z[2]) {
	B200_LOCK(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (n0 != n1) return fail(ctx, B200_ERR_INPUT_VALIDATION, "evals_0 and evals_1 must be the same length");
	if (n0 == 0) return B200_OK;
	// deferred: queue the segment; everything queued with the same challenge goes out as one launch
	// when any other entry point (or a conflicting fold) arrives
	uint8_t *w0 = (uint8_t *)e0;
	const uint8_t *r0 = (const uint8_t *)e1;
	const uint64_t bytes = n0 * 16;
	bool conflict = !ctx->pending.empty() && (ctx->pending_z[0] != z[0] || ctx->pending_z[1] != z[1] || ctx->pending.size() >= 4096);
	auto overlap = [](const uint8_t *a, uint64_t an, const uint8_t *b, uint64_t bn) { return a < b + bn && b < a + an; };
	for (size_t i = 0; i < ctx->pending.size() && !conflict; i++) {
		const b200_pending_lerp &p = ctx->pending[i];
		const uint64_t pb = p.n * 16;
		// RAW / WAW on a pending output, or WAR on a pending input
		conflict = overlap(w0, bytes, p.e0, pb) || overlap(r0, bytes, p.e0, pb) || overlap(w0, bytes, p.e1, pb);
	}
	if (conflict || !ctx->pending_fr.empty()) B200_FLUSH(ctx);
	ctx->pending_z[0] = z[0];
	ctx->pending_z[1] = z[1];
	ctx->pending.push_back(b200_pending_lerp{w0, r0, n0});
	return B200_OK;
}

}  // extern "C"
static int32_t flush_pending(b200_ctx *ctx) {
	if (!ctx->pending_fr.empty()) {
		std::vector<FrSeg> fs(ctx->pending_fr.size());
		for (size_t i = 0; i < fs.size(); i++) fs[i] = FrSeg{(const uint4 *)ctx->pending_fr[i].mat, (uint4 *)ctx->pending_fr[i].out};
		ctx->pending_fr.clear();
		const uint64_t n_out = ctx->pending_fr_n_out;
		if (fs.size() == 1) {
			k_fold_right_lut<<<grid_for(ctx, n_out, FOLD_THREADS, 2), FOLD_THREADS, LUT_BYTES + 2048, ctx->stream>>>(fs[0].mat, ctx->pending_fr_lvl, (const uint4 *)ctx->pending_fr_vec, fs[0].out, n_out);
		} else {
			void *d_segs;
			int32_t rc = stage_args(ctx, fs.data(), sizeof(FrSeg) * fs.size(), &d_segs);
			if (rc) return rc;
			k_fold_right_lut_multi<<<grid_for(ctx, n_out * fs.size(), FOLD_THREADS, 2), FOLD_THREADS, LUT_BYTES + 2048, ctx->stream>>>(
				(const FrSeg *)d_segs, (uint32_t)fs.size(), ilog2(n_out), ctx->pending_fr_lvl, (const uint4 *)ctx->pending_fr_vec);
		}
		B200_LAUNCH_CHECK(ctx);
		if (ctx->pending.empty()) return B200_OK;
	}
	std::vector<LerpSeg> segs(ctx->pending.size());
	for (size_t i = 0; i < segs.size(); i++) {
		const b200_pending_lerp &p = ctx->pending[i];
		segs[i] = LerpSeg{(uint4 *)p.e0, (const uint4 *)p.e1, p.n, p.n, make_uint4(0, 0, 0, 0), 0};
	}
	ctx->pending.clear();
	return launch_lerp(ctx, segs, ctx->pending_z);
}
extern "C" {

// Host-buffer form of the fold (what a ComputationBackend whose Vec<P> lives in host memory calls,
// hal/src/backend.rs:19-31, 65-75): chunked 4-slot pipeline  H2D(e0,e1) -> fold kernel -> D2H(e0)  on
// three streams so that both PCIe directions and the kernel overlap.  Synchronous.
int32_t b200_extrapolate_line_host(b200_ctx *ctx, void *host_e0, const void *host_e1, uint64_t n, const uint64_t z[2]) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (n == 0) return B200_OK;
	// Chunks of up to 2^20 elements (16 MiB per operand: below that the copy engines lose efficiency,
	// measured 2^18: 6.2 ms, 2^20: 5.67 ms per 2^24-coefficient fold).  The only transfers that do not
	// overlap are the first upload (pipeline fill) and the last download (drain), so the chunk size ramps
	// up from 2^17 at the start and down to 2^17 at the end.
	const uint64_t CH = 1ull << 20, CH_MIN = 1ull << 17;
	const uint32_t NS = 4;
	int32_t rc = ensure_scratch(ctx, NS * 2 * CH * 16);
	if (rc) return rc;
	if ((rc = ensure_side_streams(ctx))) return rc;
	B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	uint8_t *h0 = (uint8_t *)host_e0;
	const uint8_t *h1 = (const uint8_t *)host_e1;
	uint64_t off = 0, up = CH_MIN;
	for (uint64_t c = 0; off < n; c++) {
		const uint64_t left = n - off;
		// ramp: double from CH_MIN up to CH; near the end take half of what is left (down to CH_MIN)
		uint64_t cnt = std::min(up, CH);
		if (left <= 2 * cnt) cnt = left <= CH_MIN ? left : std::max<uint64_t>(CH_MIN, (left / 2 + 63) & ~(uint64_t)63);
		cnt = std::min(cnt, left);
		up *= 2;
		uint32_t s = (uint32_t)(c % NS);
		uint8_t *d0 = ctx->d_scratch + (uint64_t)s * 2 * CH * 16, *d1 = d0 + CH * 16;
		if (c >= NS) B200_CUDA(ctx, cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_out[s], 0));  // slot drained
		B200_CUDA(ctx, cudaMemcpyAsync(d0, h0 + off * 16, cnt * 16, cudaMemcpyHostToDevice, ctx->s_h2d));
		B200_CUDA(ctx, cudaMemcpyAsync(d1, h1 + off * 16, cnt * 16, cudaMemcpyHostToDevice, ctx->s_h2d));
		B200_CUDA(ctx, cudaEventRecord(ctx->ev_in[s], ctx->s_h2d));
		B200_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[s], 0));
		std::vector<LerpSeg> segs(1);
		segs[0] = LerpSeg{(uint4 *)d0, (const uint4 *)d1, cnt, cnt, make_uint4(0, 0, 0, 0), 0};
		if ((rc = launch_lerp(ctx, segs, z))) return rc;
		B200_CUDA(ctx, cudaEventRecord(ctx->ev_k[s], ctx->stream));
		B200_CUDA(ctx, cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_k[s], 0));
		B200_CUDA(ctx, cudaMemcpyAsync(h0 + off * 16, d0, cnt * 16, cudaMemcpyDeviceToHost, ctx->s_d2h));
		B200_CUDA(ctx, cudaEventRecord(ctx->ev_out[s], ctx->s_d2h));
		off += cnt;
	}
	B200_CUDA(ctx, cudaStreamSynchronize(ctx->s_d2h));
	B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return B200_OK;
}

int32_t b200_fold_multilinears_high_to_low(b200_ctx *ctx, const b200_dev_ptr *mls, uint32_t m, uint32_t n_vars,
										   const uint64_t *prefix, const uint64_t *suffix, const uint64_t z[2], uint64_t *new_lens) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (n_vars == 0 || n_vars > 60) return fail(ctx, B200_ERR_INPUT_VALIDATION, "n_vars must be in [1, 60]");
	uint64_t half = 1ull << (n_vars - 1);
	std::vector<LerpSeg> segs(m);
	for (uint32_t t = 0; t < m; t++) {
		uint64_t p = prefix[t];
		if (p > 2 * half) return fail(ctx, B200_ERR_INPUT_VALIDATION, "multilinear %u: prefix %llu exceeds 2^n_vars", t, (unsigned long long)p);
		uint64_t pivot = p > half ? p - half : 0;
		uint64_t upper = std::min(p, half);
		uint4 *e0 = (uint4 *)mls[t];
		segs[t] = LerpSeg{e0, e0 + half, pivot, upper, to_u4(suffix + 2 * t), 0};
		if (new_lens) new_lens[t] = upper;
	}
	return launch_lerp(ctx, segs, z);
}

int32_t b200_fold_multilinears_low_to_high(b200_ctx *ctx, const b200_dev_ptr *mls, const b200_dev_ptr *outs, uint32_t m, uint32_t n_vars,
										   const uint64_t *prefix, const uint64_t *suffix, const uint64_t z[2], uint64_t *new_lens) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (n_vars == 0 || n_vars > 60) return fail(ctx, B200_ERR_INPUT_VALIDATION, "n_vars must be in [1, 60]");
	std::vector<LerpSeg> live;
	for (uint32_t t = 0; t < m; t++) {
		uint64_t p = prefix[t];
		if (p > (1ull << n_vars)) return fail(ctx, B200_ERR_INPUT_VALIDATION, "multilinear %u: prefix %llu exceeds 2^n_vars", t, (unsigned long long)p);
		uint64_t upper = (p + 1) / 2;
		const uint8_t *in = (const uint8_t *)mls[t], *out = (const uint8_t *)outs[t];
		// fold_right_lerp is out of place (sumcheck_folding.rs:121-143); an overlap would race
		if (upper && in < out + upper * 16 && out < in + p * 16) return fail(ctx, B200_ERR_INPUT_VALIDATION, "multilinear %u: output overlaps the input", t);
		if (new_lens) new_lens[t] = upper;
		if (upper) live.push_back(LerpSeg{(uint4 *)outs[t], (const uint4 *)mls[t], p / 2, upper, to_u4(suffix + 2 * t), 0});
	}
	if (live.empty()) return B200_OK;
	bool small = ctx->tune_fold == 2;
	for (auto &sg : live) small = small && !(sg.upper >> 31);  // 32-bit tile arithmetic in the TMA kernel
	if (!small) return launch_lerp_variant<512, 2, 2, true>(ctx, live, z);
	return launch_lerp_tma<true>(ctx, live, z);
}

int32_t b200_tensor_expand(b200_ctx *ctx, b200_dev_ptr data, uint64_t data_len, uint32_t log_n, const uint64_t *coords, uint32_t k) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (log_n + k > 60 || data_len != (1ull << (log_n + k))) return fail(ctx, B200_ERR_INPUT_VALIDATION, "invalid data length: %llu", (unsigned long long)data_len);
	if (k == 0) return B200_OK;
	uint32_t k_small = 0;
	if (log_n <= EXP_SMALL_LOG && ctx->tune_expand_outer) {
		// outer-product plan (kernels.cuh): ONE k_expand_pair launch for in' (the data brought to 2^6 elements), T_mid
		// (up to 2^12 / |in'| entries) and T_hi (<= 2^10 entries); then in'' = in' x T_mid and out = in'' x T_hi
		constexpr uint32_t LOG_A = 6, LOG_HI = 10;
		const uint32_t a1 = log_n < LOG_A ? std::min(k, LOG_A - log_n) : 0, l1 = log_n + a1;  // |in'| = 2^l1
		const uint32_t k_mid = std::min(k - a1, EXP_SMALL_LOG - std::min(l1, EXP_SMALL_LOG)), l2 = l1 + k_mid;
		const uint32_t k_hi = std::min(k - a1 - k_mid, LOG_HI);
		if (!ctx->d_expand_ws) B200_CUDA(ctx, cudaMalloc((void **)&ctx->d_expand_ws, 4 * (16u << EXP_SMALL_LOG)));
		uint4 *ws_in = (uint4 *)ctx->d_expand_ws, *ws_mid = ws_in + (1u << EXP_SMALL_LOG), *ws_t1 = ws_mid + (1u << EXP_SMALL_LOG), *ws_t2 = ws_t1 + (1u << EXP_SMALL_LOG);
		ExpandPairArgs P;
		memset(&P, 0, sizeof P);
		uint32_t np = 0, cap_log = 0;
		// in': from the data; written in place when nothing follows
		P.src[np] = (const uint4 *)data, P.dst[np] = (k_mid || k_hi) ? ws_in : (uint4 *)data, P.log_n[np] = log_n, P.k[np] = a1;
		for (uint32_t i = 0; i < a1; i++) P.coords[np][i] = to_u4(coords + 2 * i);
		cap_log = std::max(cap_log, l1), np++;
		if (k_mid) {
			P.dst[np] = ws_t1, P.k[np] = k_mid;
			for (uint32_t i = 0; i < k_mid; i++) P.coords[np][i] = to_u4(coords + 2 * (a1 + i));
			cap_log = std::max(cap_log, k_mid), np++;
		}
		if (k_hi) {
			P.dst[np] = ws_t2, P.k[np] = k_hi;
			for (uint32_t i = 0; i < k_hi; i++) P.coords[np][i] = to_u4(coords + 2 * (a1 + k_mid + i));
			cap_log = std::max(cap_log, k_hi), np++;
		}
		const uint32_t pair_smem = ((FIELD_TABLE_BYTES + 127) & ~127u) + 2 * std::max(16u << cap_log, 32u * EXP_SMALL_LOG);
		k_expand_pair<<<np, 1024, pair_smem, ctx->stream>>>(ctx->d_tables, P);
		B200_LAUNCH_CHECK(ctx);
		const uint4 *cur = ws_in;
		if (k_mid) {
			uint4 *dst = k_hi ? ws_mid : (uint4 *)data;
			k_expand_outer<<<std::min<uint32_t>(1u << k_mid, (uint32_t)ctx->n_sms), 1024, LUT_BYTES + 2048 + (16u << l1), ctx->stream>>>(cur, 1u << l1, ws_t1, 1u << k_mid, dst);
			B200_LAUNCH_CHECK(ctx);
			cur = dst;
		}
		if (k_hi) {
			k_expand_outer<<<std::min<uint32_t>(1u << k_hi, (uint32_t)ctx->n_sms), 1024, LUT_BYTES + 2048 + (16u << l2), ctx->stream>>>(cur, 1u << l2, ws_t2, 1u << k_hi, (uint4 *)data);
			B200_LAUNCH_CHECK(ctx);
		}
		k_small = a1 + k_mid + k_hi;  // the rounds done so far; anything beyond 2^22 elements continues with the fused doubling rounds
	} else if (log_n <= EXP_SMALL_LOG) k_small = std::min(k, EXP_SMALL_LOG - log_n);  // rounds that fit one CTA's shared memory
	if (k_small && !(log_n <= EXP_SMALL_LOG && ctx->tune_expand_outer)) {
		std::vector<uint4> h(k_small);
		for (uint32_t i = 0; i < k_small; i++) h[i] = to_u4(coords + 2 * i);
		void *dc;
		int32_t rc = stage_args(ctx, h.data(), sizeof(uint4) * k_small, &dc);
		if (rc) return rc;
		k_expand_small<<<1, 1024, k_small * (NLUT_BYTES + 2048) + (16u << (log_n + k_small)), ctx->stream>>>((uint4 *)data, log_n, (const uint4 *)dc, k_small);
		B200_LAUNCH_CHECK(ctx);
	}
	// the rest: up to three rounds per launch
	for (uint32_t r = k_small; r < k;) {
		// the short launch goes first: the last (largest) launches fuse three rounds
		const uint32_t R = (k - r) % 3 ? (k - r) % 3 : 3;
		ExpandArgs A;
		A.data = (uint4 *)data;
		A.n0 = 1ull << (log_n + r);
		for (uint32_t t = 0; t < 3; t++) A.r[t] = t < R ? to_u4(coords + 2 * (r + t)) : make_uint4(0, 0, 0, 0);
		uint32_t grid = (uint32_t)std::min<uint64_t>((A.n0 + EXP_THREADS - 1) / EXP_THREADS, (uint64_t)ctx->n_sms);
		if (R == 3) k_expand_k64<3><<<grid, EXP_THREADS, 3 * LUT_BYTES + 6144, ctx->stream>>>(A);
		else if (R == 2) k_expand_k64<2><<<grid, EXP_THREADS, 2 * LUT_BYTES + 6144, ctx->stream>>>(A);
		else k_expand_k64<1><<<grid, EXP_THREADS, 1 * LUT_BYTES + 6144, ctx->stream>>>(A);
		B200_LAUNCH_CHECK(ctx);
		r += R;
	}
	return B200_OK;
}

int32_t b200_tensor_product_full_query(b200_ctx *ctx, const uint64_t *query, uint32_t k, b200_dev_ptr out, uint64_t n_out) {
	B200_LOCK(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (k > 60 || n_out != (1ull << k)) return fail(ctx, B200_ERR_INPUT_VALIDATION, "output must hold 2^%u elements", k);
	const uint64_t one[2] = {1, 0};
	int32_t rc = b200_fill(ctx, out, 1, one);
	if (rc) return rc;
	return b200_tensor_expand(ctx, out, n_out, 0, query, k);
}

static bool valid_level(uint32_t lvl) { return lvl == 0 || (lvl >= 3 && lvl <= 7); }

int32_t b200_inner_product(b200_ctx *ctx, b200_dev_ptr a, uint64_t n_a, uint32_t lvl, b200_dev_ptr b, uint64_t n_b, uint32_t *slot) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx || !slot) return B200_ERR_INPUT_VALIDATION;
	if (lvl > 7 || (n_a << (7 - lvl)) != n_b) return fail(ctx, B200_ERR_INPUT_VALIDATION, "invalid input: a_edeg=%u |a|=%llu |b|=%llu", lvl, (unsigned long long)n_a, (unsigned long long)n_b);
	if (!valid_level(lvl)) return fail(ctx, B200_ERR_INPUT_VALIDATION, "unsupported tower level %u", lvl);
	int32_t rc = new_slot(ctx, slot);
	if (rc) return rc;
	if (n_b == 0) return B200_OK;
	if (lvl == 7 && n_b >= 4096 && n_b % tc::CHUNK == 0) {
		// B128 x B128: one inner-product job on the tensor cores (roundevals_tc.cuh)
		// (the kernel works on job pairs: the two halves of the range form the pair when they stay chunk-aligned)
		const bool split = (n_b / 2) % tc::CHUNK == 0;
		const uint64_t len = split ? n_b / 2 : n_b;
		std::vector<tc::TcJob> jobs(1, tc::TcJob{(const uint4 *)a, nullptr, (const uint4 *)b, nullptr});
		std::vector<tc::TcTarget> targets(1, tc::TcTarget{0, *slot, make_uint4(1, 0, 0, 0)});
		if (split) {
			jobs.push_back(tc::TcJob{(const uint4 *)a + len, nullptr, (const uint4 *)b + len, nullptr});
			targets.push_back(tc::TcTarget{1, *slot, make_uint4(1, 0, 0, 0)});
		}
		return launch_tc_pairs(ctx, jobs, len, targets);
	}
	k_inner_product<<<grid_for(ctx, n_b, 256, 2), 256, FIELD_TABLE_BYTES, ctx->stream>>>(ctx->d_tables, (const uint4 *)a, lvl, (const uint4 *)b, n_b, ctx->d_results + *slot);
	B200_LAUNCH_CHECK(ctx);
	return B200_OK;
}

static int32_t fold_mat(b200_ctx *ctx, bool right, b200_dev_ptr mat, uint64_t n_mat, uint32_t lvl, b200_dev_ptr vec, uint64_t n_vec, b200_dev_ptr out, uint64_t n_out) {
	B200_LOCK(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	// the byte-LUT fold_right is deferred like extrapolate_line: calls with the same query on disjoint buffers are batched
	const bool lut_right = right && valid_level(lvl) && is_pow2(n_mat) && is_pow2(n_vec) && (n_vec << lvl) == 128 && n_out >= 1024 && n_out == n_mat;
	if (lut_right && !ctx->pending_fr.empty() && ctx->pending.empty() && !ctx->tail_active && ctx->pending_fr_vec == vec && ctx->pending_fr_lvl == lvl &&
		ctx->pending_fr_n_out == n_out && ctx->pending_fr.size() < 4096) {
		const uint8_t *m8 = (const uint8_t *)mat, *o8 = (const uint8_t *)out, *v8 = (const uint8_t *)vec;
		const uint64_t bytes = n_out * 16;
		auto overlap = [](const uint8_t *a, uint64_t an, const uint8_t *b, uint64_t bn) { return a < b + bn && b < a + an; };
		bool conflict = overlap(o8, bytes, v8, n_vec * 16) || (m8 != o8 && overlap(m8, bytes, o8, bytes));
		for (size_t i = 0; i < ctx->pending_fr.size() && !conflict; i++) {
			const b200_pending_fold_right &p = ctx->pending_fr[i];
			conflict = overlap(o8, bytes, p.out, bytes) || overlap(m8, bytes, p.out, bytes) || overlap(o8, bytes, p.mat, bytes);
		}
		if (!conflict) {
			ctx->pending_fr.push_back(b200_pending_fold_right{m8, (uint8_t *)out});
			return B200_OK;
		}
	}
	B200_FLUSH(ctx);
	if (lvl > 7) return fail(ctx, B200_ERR_INPUT_VALIDATION, "invalid evals: tower_level=%u > 7", lvl);
	if (!valid_level(lvl)) return fail(ctx, B200_ERR_INPUT_VALIDATION, "unsupported tower level %u", lvl);
	if (!is_pow2(n_mat)) return fail(ctx, B200_ERR_INPUT_VALIDATION, "the length of `mat` must be a power of 2");
	if (!is_pow2(n_vec)) return fail(ctx, B200_ERR_INPUT_VALIDATION, "the length of `vec` must be a power of 2");
	uint32_t log_evals = ilog2(n_mat) + 7 - lvl, log_q = ilog2(n_vec);
	if (log_q > log_evals) return fail(ctx, B200_ERR_INPUT_VALIDATION, "query larger than evals");
	uint64_t expect_out = 1ull << (log_evals - log_q);
	if (n_out != expect_out) return fail(ctx, B200_ERR_INPUT_VALIDATION, "output has %llu elements, expected %llu", (unsigned long long)n_out, (unsigned long long)expect_out);
	if (right && (n_vec << lvl) == 128 && n_out >= 1024) {
		const uint8_t *m8 = (const uint8_t *)mat, *o8 = (const uint8_t *)out, *v8 = (const uint8_t *)vec;
		const bool self_ok = !(o8 < v8 + n_vec * 16 && v8 < o8 + n_out * 16) && (m8 == o8 || !(m8 < o8 + n_out * 16 && o8 < m8 + n_out * 16));
		if (self_ok) {  // first of a possible batch: launched by the next other entry point (or a conflicting fold)
			ctx->pending_fr.push_back(b200_pending_fold_right{m8, (uint8_t *)out});
			ctx->pending_fr_vec = vec, ctx->pending_fr_lvl = lvl, ctx->pending_fr_n_out = n_out;
			return B200_OK;
		}
		k_fold_right_lut<<<grid_for(ctx, n_out, FOLD_THREADS, 2), FOLD_THREADS, LUT_BYTES + 2048, ctx->stream>>>((const uint4 *)mat, lvl, (const uint4 *)vec, (uint4 *)out, n_out);
		B200_LAUNCH_CHECK(ctx);
		return B200_OK;
	}
	if (!right && lvl == 0 && n_out == 128 && n_vec >= 8192 && n_vec % (2 * tc::CHUNK) == 0 && ctx->tune_round_evals_tc) {
		// bit-packed matrix, LONG query, 128 outputs (evaluate_partial_high down to 7 variables: ring-switch partial
		// evaluations): out[q] = sum_j vec[j] * bit_q(word_j) = an outer-product bit-GEMM on the tensor cores
		const uint64_t len = n_vec / 2;
		std::vector<tc::TcJob> jobs{tc::TcJob{(const uint4 *)vec, nullptr, (const uint4 *)mat, nullptr},
									tc::TcJob{(const uint4 *)vec + len, nullptr, (const uint4 *)mat + len, nullptr}};
		uint32_t *gmat = nullptr;
		int32_t rc = launch_tc_pairs(ctx, jobs, len, {}, 0, &gmat);
		if (rc) return rc;
		tc::k_tc_outer_combine<<<1, 128, 0, ctx->stream>>>(gmat, 2, (uint4 *)out);
		B200_LAUNCH_CHECK(ctx);
		return B200_OK;
	}
	if (!right && lvl == 0 && n_vec <= 128 && n_out >= 4096 && n_out % 128 == 0) {
		// bit-packed matrix, short query: warp-transposed bit blocks + byte-LUT linear map
		const uint64_t n_warps = n_out / 128;
		uint32_t g = (uint32_t)std::min<uint64_t>((n_warps * 32 + FOLD_THREADS - 1) / FOLD_THREADS, (uint64_t)ctx->n_sms * 2);
		k_fold_left_b1_lut<<<g, FOLD_THREADS, LUT_BYTES + 2048, ctx->stream>>>((const uint4 *)mat, (const uint4 *)vec, (uint32_t)n_vec, (uint4 *)out, n_out);
		B200_LAUNCH_CHECK(ctx);
		return B200_OK;
	}
	uint32_t grid = grid_for(ctx, n_out, 256, 2);
	if (right) k_fold_mat<true><<<grid, 256, FIELD_TABLE_BYTES, ctx->stream>>>(ctx->d_tables, (const uint4 *)mat, lvl, (const uint4 *)vec, log_q, (uint4 *)out, n_out);
	else k_fold_mat<false><<<grid, 256, FIELD_TABLE_BYTES, ctx->stream>>>(ctx->d_tables, (const uint4 *)mat, lvl, (const uint4 *)vec, log_q, (uint4 *)out, n_out);
	B200_LAUNCH_CHECK(ctx);
	return B200_OK;
}
int32_t b200_fold_left(b200_ctx *ctx, b200_dev_ptr mat, uint64_t n_mat, uint32_t lvl, b200_dev_ptr vec, uint64_t n_vec, b200_dev_ptr out, uint64_t n_out) {
	return fold_mat(ctx, false, mat, n_mat, lvl, vec, n_vec, out, n_out);
}
int32_t b200_fold_right(b200_ctx *ctx, b200_dev_ptr mat, uint64_t n_mat, uint32_t lvl, b200_dev_ptr vec, uint64_t n_vec, b200_dev_ptr out, uint64_t n_out) {
	return fold_mat(ctx, true, mat, n_mat, lvl, vec, n_vec, out, n_out);
}

// ---- Groestl-256 Merkle commitments on device-resident data ----------------------------------------------
int32_t b200_groestl256_leaves(b200_ctx *ctx, b200_dev_ptr data, uint64_t n_leaves, uint64_t leaf_elems, b200_dev_ptr digests) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (leaf_elems == 0 || leaf_elems > (1u << 24)) return fail(ctx, B200_ERR_INPUT_VALIDATION, "IncorrectBatchSize: leaves of %llu elements", (unsigned long long)leaf_elems);
	if (n_leaves == 0) return B200_OK;
	groestl::k_groestl_leaves<<<grid_for(ctx, n_leaves, groestl::THREADS, 1), groestl::THREADS, groestl::SMEM, ctx->stream>>>(
		(const uint2 *)ctx->d_groestl_t0, (const uint4 *)data, n_leaves, (uint32_t)(leaf_elems * 16), (uint4 *)digests);
	B200_LAUNCH_CHECK(ctx);
	return B200_OK;
}
int32_t b200_groestl256_compress_pairs(b200_ctx *ctx, b200_dev_ptr in_digests, uint64_t n_pairs, b200_dev_ptr out_digests) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (n_pairs == 0) return B200_OK;
	groestl::k_groestl_compress_pairs<<<grid_for(ctx, n_pairs, groestl::THREADS, 1), groestl::THREADS, groestl::SMEM, ctx->stream>>>(
		(const uint2 *)ctx->d_groestl_t0, (const uint4 *)in_digests, n_pairs, (uint4 *)out_digests);
	B200_LAUNCH_CHECK(ctx);
	return B200_OK;
}
int32_t b200_merkle_build(b200_ctx *ctx, b200_dev_ptr elements, uint64_t n_elems, uint64_t batch_size, b200_dev_ptr nodes, uint64_t n_nodes) {
	B200_LOCK(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (batch_size == 0 || n_elems % batch_size) return fail(ctx, B200_ERR_INPUT_VALIDATION, "IncorrectBatchSize");
	const uint64_t n_leaves = n_elems / batch_size;
	if (!is_pow2(n_leaves)) return fail(ctx, B200_ERR_INPUT_VALIDATION, "PowerOfTwoLengthRequired");
	if (n_nodes != 2 * n_leaves - 1) return fail(ctx, B200_ERR_INPUT_VALIDATION, "IncorrectVectorLen: the tree has %llu nodes", (unsigned long long)(2 * n_leaves - 1));
	int32_t rc = b200_groestl256_leaves(ctx, elements, n_leaves, batch_size, nodes);
	uint8_t *prev = (uint8_t *)nodes, *cur = prev + 32 * n_leaves;
	for (uint64_t n = n_leaves / 2; n >= 1 && !rc; n /= 2) {
		rc = b200_groestl256_compress_pairs(ctx, prev, n, cur);
		prev = cur;
		cur += 32 * n;
	}
	return rc;
}

// ---- GF(2)-linear maps on B128 (basis changes) ------------------------------------------------------
int32_t b200_linear_map(b200_ctx *ctx, b200_dev_ptr src, b200_dev_ptr dst, uint64_t n, const uint64_t *basis_images) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx || !basis_images) return B200_ERR_INPUT_VALIDATION;
	if (n == 0) return B200_OK;
	const uint8_t *a = (const uint8_t *)src, *b = (const uint8_t *)dst;
	if (a != b && a < b + n * 16 && b < a + n * 16) return fail(ctx, B200_ERR_INPUT_VALIDATION, "linear_map: partially overlapping source and destination");
	void *dw;
	int32_t rc = stage_args(ctx, basis_images, 128 * 16, &dw);
	if (rc) return rc;
	k_linear_map<<<grid_for(ctx, n, FOLD_THREADS, 2), FOLD_THREADS, LUT_BYTES + 2048, ctx->stream>>>((const uint4 *)src, (const uint4 *)dw, (uint4 *)dst, n);
	B200_LAUNCH_CHECK(ctx);
	return B200_OK;
}

// BinaryField128bPolyval (field/src/polyval.rs): GF(2)[X] / (X^128 + X^127 + X^126 + X^121 + 1), elements stored in
// Montgomery form; product = a * b * X^-128 (arch/portable/packed_polyval_128.rs:88-122).  Host-side, bit-serial: used
// for one-off scalar work only (deriving the basis change, converting round values).
void b200_host_polyval_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]) {
	const hostf::u128 p = hostf::polyval_mul(hostf::from_words(a), hostf::from_words(b));
	out[0] = (uint64_t)p;
	out[1] = (uint64_t)(p >> 64);
}
// The tower <-> POLYVAL isomorphism as 128 basis images each way (images[k] = phi(beta_k), 2 words per image),
// DERIVED from the two published multiplicative generators (binary_field.rs:747, polyval.rs:496): phi is the field
// isomorphism with phi(g_tower) = g_polyval, so phi(sum_k c_k g^k) = sum_k c_k g'^k; the coordinates of beta_i in the
// power basis of g come from inverting a 128 x 128 GF(2) matrix.  tests/ compare the result with the reference's
// BINARY_TO_POLYVAL_TRANSFORMATION / POLYVAL_TO_BINARY_TRANSFORMATION tables (polyval.rs:516-788).
int32_t b200_host_polyval_basis_change(uint64_t tower_to_polyval[256], uint64_t polyval_to_tower[256]) {
	using hostf::u128;
	const u128 g_t = ((u128)0x2E895399AF449ACEull << 64) | 0x499596F6E5FCCAFAull;  // BinaryField128b::MULTIPLICATIVE_GENERATOR
	const u128 g_p = ((u128)0x072bdf2504ce49c0ull << 64) | 0x3105433c1c25a4a7ull;  // BinaryField128bPolyval::MULTIPLICATIVE_GENERATOR (stored form)
	const u128 one_p = ((u128)0xc200000000000000ull << 64) | 1ull;                 // BinaryField128bPolyval::ONE (X^128 mod p)
	u128 pw_t[128], pw_p[128];
	pw_t[0] = 1, pw_p[0] = one_p;
	for (int k = 1; k < 128; k++) {
		pw_t[k] = hostf::mul128(pw_t[k - 1], g_t);
		pw_p[k] = hostf::polyval_mul(pw_p[k - 1], g_p);
	}
	auto solve = [](const u128 *from, const u128 *to, uint64_t *out) -> bool {
		// rows: (from[k] | to[k]); Gaussian elimination brings `from` to the identity, `to` then holds the images
		u128 a[128], b[128];
		for (int k = 0; k < 128; k++) a[k] = from[k], b[k] = to[k];
		for (int col = 0; col < 128; col++) {
			int piv = -1;
			for (int r = col; r < 128; r++)
				if ((a[r] >> col) & 1) {
					piv = r;
					break;
				}
			if (piv < 0) return false;
			std::swap(a[piv], a[col]);
			std::swap(b[piv], b[col]);
			for (int r = 0; r < 128; r++)
				if (r != col && ((a[r] >> col) & 1)) a[r] ^= a[col], b[r] ^= b[col];
		}
		for (int k = 0; k < 128; k++) out[2 * k] = (uint64_t)b[k], out[2 * k + 1] = (uint64_t)(b[k] >> 64);
		return true;
	};
	if (!solve(pw_t, pw_p, tower_to_polyval) || !solve(pw_p, pw_t, polyval_to_tower)) return B200_ERR_INPUT_VALIDATION;
	return B200_OK;
}

// ---- expressions --------------------------------------------------------------------------------
int32_t b200_expr_compile(b200_ctx *ctx, const b200_expr_step *steps, uint32_t n_steps, b200_expr **out) {
	B200_LOCK(ctx);
	if (!ctx || !out) return B200_ERR_INPUT_VALIDATION;
	if (n_steps > MAX_EXPR_STEPS) return fail(ctx, B200_ERR_INPUT_VALIDATION, "expression has %u steps (max %u)", n_steps, MAX_EXPR_STEPS);
	uint32_t n_vars = 0;
	for (uint32_t s = 0; s < n_steps; s++) {
		const b200_expr_step &st = steps[s];
		switch (st.op) {
		case 0:
		case 1:
			if (st.l >= s || st.r >= s) return fail(ctx, B200_ERR_INPUT_VALIDATION, "step %u references a later step", s);
			break;
		case 2:
			if (st.l >= s) return fail(ctx, B200_ERR_INPUT_VALIDATION, "step %u references a later step", s);
			break;
		case 3: break;
		case 4: n_vars = std::max(n_vars, st.l + 1); break;
		default: return fail(ctx, B200_ERR_INPUT_VALIDATION, "step %u: unknown op %u", s, st.op);
		}
	}
	std::unique_ptr<b200_expr> e(new b200_expr);
	e->ctx = ctx;
	e->steps.assign(steps, steps + n_steps);
	e->n_vars = n_vars;
	e->poly_ok = plan::expand(steps, n_steps, 2, 64, e->poly);
	static std::atomic<uint64_t> next_uid{1};
	e->uid = next_uid.fetch_add(1);
	*out = e.release();
	return B200_OK;
}
static std::mutex g_slab_mu;
void b200_expr_free(b200_expr *e) {
	if (!e) return;
	if (e->slab) {
		bool last;
		{
			std::lock_guard<std::mutex> g(g_slab_mu);
			last = --e->slab->refs == 0;
		}
		if (last) {
			ctx_free(e->ctx, e->slab->base);
			delete e->slab;
		}
	} else
		ctx_free(e->ctx, e->d_steps);
	delete e;
}
uint32_t b200_expr_n_vars(const b200_expr *e) { return e ? e->n_vars : 0; }

// The device copy of the steps is made on first use by an interpreter kernel: compile_expr itself stays a host-side
// operation (the prover compiles its compositions every round, bivariate_product.rs:311-314; the traced and monomial
// paths never read the steps on the device), so it must not cost a cudaMalloc.
static DevExpr dev_expr(const b200_expr *e) {
	b200_expr *m = const_cast<b200_expr *>(e);
	if (!m->d_steps && !m->steps.empty()) {
		if (cudaMalloc(&m->d_steps, sizeof(b200_expr_step) * m->steps.size()) != cudaSuccess ||
			cudaMemcpy(m->d_steps, m->steps.data(), sizeof(b200_expr_step) * m->steps.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
			cudaGetLastError();
			if (m->d_steps) cudaFree(m->d_steps);
			m->d_steps = nullptr;
		}
	}
	return DevExpr{m->d_steps, (uint32_t)m->steps.size(), m->n_vars};
}
// First use of MANY expressions by one call (75 compositions + 75 leading terms of a keccak zerocheck round: 150 x
// (cudaMalloc + synchronous copy) = 1.1 ms): one allocation, one copy.  Failures are left to dev_expr / B200_EXPR_READY.
static void upload_exprs(std::initializer_list<std::pair<const b200_expr *const *, uint32_t>> lists) {
	std::vector<b200_expr *> todo;
	size_t n_steps = 0;
	for (auto &l : lists)
		for (uint32_t c = 0; l.first && c < l.second; c++) {
			b200_expr *m = const_cast<b200_expr *>(l.first[c]);
			if (m && !m->d_steps && !m->steps.empty() && std::find(todo.begin(), todo.end(), m) == todo.end()) {
				todo.push_back(m);
				n_steps += m->steps.size();
			}
		}
	if (todo.size() < 2) return;
	std::vector<b200_expr_step> host;
	host.reserve(n_steps);
	for (b200_expr *m : todo) host.insert(host.end(), m->steps.begin(), m->steps.end());
	void *base = nullptr;
	if (cudaMalloc(&base, sizeof(b200_expr_step) * n_steps) != cudaSuccess || cudaMemcpy(base, host.data(), sizeof(b200_expr_step) * n_steps, cudaMemcpyHostToDevice) != cudaSuccess) {
		cudaGetLastError();
		if (base) cudaFree(base);
		return;
	}
	StepSlab *slab = new StepSlab{base, (uint32_t)todo.size()};
	size_t off = 0;
	for (b200_expr *m : todo) {
		m->d_steps = reinterpret_cast<b200_expr_step *>(base) + off;
		m->slab = slab;
		off += m->steps.size();
	}
}
#define B200_EXPR_READY(ctx, de) \
	if ((de).n_steps && !(de).steps) return fail(ctx, B200_ERR_ALLOC, "out of device memory (expression steps)")

int32_t b200_compute_composite(b200_ctx *ctx, const b200_dev_ptr *inputs, uint32_t n_inputs, uint64_t row_len, b200_dev_ptr out, uint64_t n_out, const b200_expr *expr) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx || !expr) return B200_ERR_INPUT_VALIDATION;
	if (row_len != n_out) return fail(ctx, B200_ERR_INPUT_VALIDATION, "inputs and output must be the same length");
	if (expr->n_vars > n_inputs) return fail(ctx, B200_ERR_INPUT_VALIDATION, "composition not match with input");
	if (n_out == 0) return B200_OK;
	void *dptrs = nullptr;
	if (n_inputs) {
		int32_t rc = stage_args(ctx, inputs, sizeof(void *) * n_inputs, &dptrs);
		if (rc) return rc;
	}
	const DevExpr de = dev_expr(expr);
	B200_EXPR_READY(ctx, de);
	k_compute_composite<<<grid_for(ctx, n_out, 256, 2), 256, FIELD_TABLE_BYTES, ctx->stream>>>(ctx->d_tables, PtrList{(const uint4 *const *)dptrs, n_inputs}, de, (uint4 *)out, n_out);
	B200_LAUNCH_CHECK(ctx);
	return B200_OK;
}

int32_t b200_pairwise_product_reduce(b200_ctx *ctx, b200_dev_ptr input, uint64_t n_in, const b200_dev_ptr *outs, const uint64_t *lens, uint32_t n_rounds) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (!is_pow2(n_in)) return fail(ctx, B200_ERR_INPUT_VALIDATION, "input length must be a power of 2: %llu", (unsigned long long)n_in);
	if (n_in < 2) return fail(ctx, B200_ERR_INPUT_VALIDATION, "input length must be greater than or equal to 2 in order to perform at least one reduction: %llu", (unsigned long long)n_in);
	uint32_t log_n = ilog2(n_in);
	if (n_rounds != log_n) return fail(ctx, B200_ERR_INPUT_VALIDATION, "round_outputs.len() does not match the expected length: %u != %u", n_rounds, log_n);
	for (uint32_t r = 0; r < n_rounds; r++)
		if (lens[r] != (1ull << (log_n - r - 1))) return fail(ctx, B200_ERR_INPUT_VALIDATION, "round_outputs[%u].len() = %llu, expected %llu", r, (unsigned long long)lens[r], 1ull << (log_n - r - 1));
	const uint4 *src = (const uint4 *)input;
	for (uint32_t r = 0; r < n_rounds; r++) {
		k_pairwise_product<<<grid_for(ctx, lens[r], 256, 2), 256, FIELD_TABLE_BYTES, ctx->stream>>>(ctx->d_tables, src, (uint4 *)outs[r], lens[r]);
		B200_LAUNCH_CHECK(ctx);
		src = (const uint4 *)outs[r];
	}
	return B200_OK;
}

// ---- KernelExecutor ------------------------------------------------------------------------------
// Outside a kernel scope every op launches at once.  Inside one (b200_kernel_scope_begin .. _end, what the
// layer's accumulate_kernels / map_kernels wrap around the caller's closure, layer.rs:134-245) the ops are
// RECORDED and lowered together at scope end: the only closure the reference prover issues
// (core/src/protocols/sumcheck/v3/bivariate_product.rs:343-405: sums of Var*Var over the high halves, add(lo, hi)
// into the Local buffers, sums of Var*Var over the Locals) becomes inner-product jobs of the tensor-core kernel
// with the (lo, hi) pointer pairs as operands -- the Local buffers are never written.
}  // extern "C"
static int32_t kernel_decl_now(b200_ctx *ctx, const uint64_t init[2], uint32_t slot) {
	if (init[0] | init[1]) {
		k_set_slot<<<1, 1, 0, ctx->stream>>>(ctx->d_results + slot, to_u4(init));
		B200_LAUNCH_CHECK(ctx);
	}
	return B200_OK;
}
static int32_t kernel_sum_now(b200_ctx *ctx, const b200_dev_ptr *inputs, uint32_t n_inputs, uint64_t row_len, const b200_expr *expr, const uint64_t coeff[2], uint32_t slot) {
	if (row_len == 0) return B200_OK;
	void *dptrs = nullptr;
	if (n_inputs) {
		int32_t rc = stage_args(ctx, inputs, sizeof(void *) * n_inputs, &dptrs);
		if (rc) return rc;
	}
	const DevExpr de = dev_expr(expr);
	B200_EXPR_READY(ctx, de);
	k_sum_composition<<<grid_for(ctx, row_len, 256, 2), 256, FIELD_TABLE_BYTES, ctx->stream>>>(ctx->d_tables, PtrList{(const uint4 *const *)dptrs, n_inputs}, de, row_len, to_u4(coeff), ctx->d_results + slot);
	B200_LAUNCH_CHECK(ctx);
	return B200_OK;
}
static int32_t kernel_add_now(b200_ctx *ctx, uint32_t log_len, const void *a, const void *b, void *dst) {
	uint64_t n = 1ull << log_len;
	k_add<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>((const uint4 *)a, (const uint4 *)b, (uint4 *)dst, n);
	B200_LAUNCH_CHECK(ctx);
	return B200_OK;
}

// Lower the recorded ops of a scope.  Fused when every sum is a polynomial of degree <= 2 over rows of one length
// (>= 4096, a multiple of the tensor-core chunk), every add writes a whole Local exactly once from two non-Local
// sources before any sum reads it, and sums read only such Locals or buffers the scope never writes.
static int32_t run_trace(b200_ctx *ctx) {
	std::vector<b200_trace_op> &ops = ctx->trace;
	// Locals sorted by address: membership of a pointer in O(log n) (a PIOP round maps 200 Locals and issues 600 ops)
	std::vector<std::pair<const uint8_t *, int>> order(ctx->locals.size());
	for (size_t i = 0; i < order.size(); i++) order[i] = {ctx->locals[i].first, (int)i};
	std::sort(order.begin(), order.end());
	auto local_ix = [&](const void *p) -> int {
		auto it = std::upper_bound(order.begin(), order.end(), std::make_pair((const uint8_t *)p, INT32_MAX));
		if (it == order.begin()) return -1;
		--it;
		return (const uint8_t *)p < it->first + ctx->locals[it->second].second ? it->second : -1;
	};
	bool ok = ctx->tune_round_evals_tc != 0;
	uint64_t len = 0;
	std::map<const void *, std::pair<const void *, const void *>> defs;
	for (size_t i = 0; i < ops.size() && ok; i++) {
		const b200_trace_op &op = ops[i];
		if (op.kind == 2) {
			const int li = local_ix(op.dst);
			ok = li >= 0 && op.dst == ctx->locals[li].first && (16ull << op.log_len) == ctx->locals[li].second && local_ix(op.a) < 0 && local_ix(op.b) < 0 && !defs.count(op.dst);
			if (ok) defs[op.dst] = {op.a, op.b};
		} else if (op.kind == 1) {
			ok = op.expr->poly_ok && (len == 0 || len == op.row_len);
			len = op.row_len;
			for (auto &t : op.expr->poly)
				for (uint32_t v : t.first) {
					const void *p = op.inputs[v];
					if (local_ix(p) >= 0 && !(defs.count(p) && (16ull * op.row_len) == ctx->locals[local_ix(p)].second)) ok = false;
				}
		}
	}
	const bool small = len < 4096 || len % tc::CHUNK != 0;  // below the tensor-core kernel's granularity: k_jobs_small
	ok = ok && len >= 1;
	if (!ok) {
		// in order, as recorded; Locals are zero-initialised (layer.rs:617-644)
		for (auto &l : ctx->locals) B200_CUDA(ctx, cudaMemsetAsync(l.first, 0, l.second, ctx->stream));
		for (const b200_trace_op &op : ops) {
			int32_t rc = op.kind == 0   ? kernel_decl_now(ctx, op.init, op.slot)
						 : op.kind == 1 ? kernel_sum_now(ctx, (const b200_dev_ptr *)op.inputs.data(), (uint32_t)op.inputs.size(), op.row_len, op.expr, op.coeff, op.slot)
										: kernel_add_now(ctx, op.log_len, op.a, op.b, op.dst);
			if (rc) return rc;
		}
		return B200_OK;
	}
	std::vector<tc::TcJob> jobs;
	std::vector<tc::TcTarget> targets;
	std::vector<std::pair<uint32_t, uint4>> const_terms;
	std::map<std::tuple<const void *, const void *, const void *, const void *>, uint32_t> job_ix;
	bool need_ones = false;
	for (const b200_trace_op &op : ops)
		if (op.kind == 1)
			for (auto &t : op.expr->poly) need_ones = need_ones || t.first.size() == 1;
	const uint64_t off_g = need_ones ? len * 16 : 0;
	{
		size_t n_jobs = 1;
		for (const b200_trace_op &op : ops)
			if (op.kind == 1) n_jobs += op.expr->poly.size();
		int32_t rc = ensure_scratch(ctx, off_g + (uint64_t)(n_jobs + 1) * 512 * 4);
		if (rc) return rc;
	}
	const uint4 *ones = (const uint4 *)ctx->d_scratch;
	if (need_ones) {
		k_fill<<<grid_for(ctx, len, 256, 8), 256, 0, ctx->stream>>>((uint4 *)ctx->d_scratch, len, make_uint4(1, 0, 0, 0));
		B200_LAUNCH_CHECK(ctx);
	}
	for (const b200_trace_op &op : ops) {
		if (op.kind == 0) {
			int32_t rc = kernel_decl_now(ctx, op.init, op.slot);
			if (rc) return rc;
		}
		if (op.kind != 1) continue;
		const hostf::u128 cf = hostf::from_words(op.coeff);
		for (auto &t : op.expr->poly) {
			if (t.first.empty()) {  // a constant c summed over len points: c if len is odd, else 0
				if (len & 1) {
					const hostf::u128 wgt = hostf::mul128(cf, t.second);
					uint64_t w[2] = {(uint64_t)wgt, (uint64_t)(wgt >> 64)};
					const_terms.push_back({op.slot, to_u4(w)});
				}
				continue;
			}
			const void *o[4] = {nullptr, nullptr, need_ones ? (const void *)ones : nullptr, nullptr};
			for (size_t f = 0; f < t.first.size(); f++) {
				const void *p = op.inputs[t.first[f]];
				auto d = defs.find(p);
				o[2 * f] = d == defs.end() ? p : d->second.first;
				o[2 * f + 1] = d == defs.end() ? nullptr : d->second.second;
			}
			auto key = std::make_tuple(o[0], o[1], o[2], o[3]);
			auto it = job_ix.find(key);
			if (it == job_ix.end()) {
				it = job_ix.emplace(key, (uint32_t)jobs.size()).first;
				jobs.push_back(tc::TcJob{(const uint4 *)o[0], (const uint4 *)o[1], (const uint4 *)o[2], (const uint4 *)o[3]});
			}
			const hostf::u128 wgt = t.second == 1 ? cf : hostf::mul128(cf, t.second);  // (a host-side tower product is ~0.3 us)
			uint64_t w[2] = {(uint64_t)wgt, (uint64_t)(wgt >> 64)};
			if (wgt) targets.push_back(tc::TcTarget{it->second, op.slot, to_u4(w)});
		}
	}
	for (auto &ct : const_terms) {
		k_xor_slot<<<1, 1, 0, ctx->stream>>>(ctx->d_results + ct.first, ct.second);
		B200_LAUNCH_CHECK(ctx);
	}
	if (targets.empty()) return B200_OK;
	if (small) {
		ArgPack pack;
		size_t o_j = pack.add(jobs.data(), sizeof(tc::TcJob) * jobs.size()), o_t = pack.add(targets.data(), sizeof(tc::TcTarget) * targets.size());
		uint8_t *dbase;
		int32_t rc = pack.commit(ctx, &dbase);
		if (rc) return rc;
		tc::k_jobs_small<<<(uint32_t)targets.size(), 256, FIELD_TABLE_BYTES, ctx->stream>>>(ctx->d_tables, (const tc::TcJob *)(dbase + o_j), (const tc::TcTarget *)(dbase + o_t), len, ctx->d_results);
		B200_LAUNCH_CHECK(ctx);
		return B200_OK;
	}
	return launch_tc_pairs(ctx, jobs, len, targets, off_g);
}
extern "C" {

int32_t b200_kernel_scope_begin(b200_ctx *ctx) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (ctx->tracing) return fail(ctx, B200_ERR_INPUT_VALIDATION, "kernel scopes do not nest");
	ctx->mu.lock();  // held until b200_kernel_scope_end: the scope belongs to the calling thread
	ctx->tracing = true;
	ctx->trace.clear();
	ctx->locals.clear();
	for (auto &c : ctx->local_pool) c.used = 0;
	return B200_OK;
}
int32_t b200_kernel_local(b200_ctx *ctx, uint32_t log_size, b200_dev_ptr *out) {
	B200_LOCK_REC(ctx);
	if (!ctx || !out) return B200_ERR_INPUT_VALIDATION;
	if (!ctx->tracing) return fail(ctx, B200_ERR_INPUT_VALIDATION, "b200_kernel_local outside a kernel scope");
	if (log_size > 40) return fail(ctx, B200_ERR_INPUT_VALIDATION, "Local buffer too large");
	const uint64_t bytes = std::max<uint64_t>(16ull << log_size, 256);
	for (auto &c : ctx->local_pool)
		if (c.bytes - c.used >= bytes) {
			*out = c.p + c.used;
			ctx->locals.push_back({c.p + c.used, 16ull << log_size});
			c.used += bytes;
			return B200_OK;
		}
	// chunks are never moved or freed before the context dies: pointers handed out earlier stay valid
	const uint64_t want = std::max<uint64_t>(bytes, 64ull << 20);
	CtxGuard device_now(ctx);  // (the pool grows: this path does need the context's device)
	void *p = nullptr;
	if (cudaMalloc(&p, want) != cudaSuccess) {
		cudaGetLastError();
		if (want == bytes || cudaMalloc(&p, bytes) != cudaSuccess) {
			cudaGetLastError();
			return fail(ctx, B200_ERR_ALLOC, "out of device memory (kernel-local scratch of %llu bytes)", (unsigned long long)bytes);
		}
		ctx->local_pool.push_back(b200_local_chunk{(uint8_t *)p, bytes, bytes});
	} else {
		ctx->local_pool.push_back(b200_local_chunk{(uint8_t *)p, want, bytes});
	}
	*out = p;
	ctx->locals.push_back({(uint8_t *)p, 16ull << log_size});
	return B200_OK;
}
int32_t b200_kernel_scope_end(b200_ctx *ctx) {
	B200_LOCK(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	if (!ctx->tracing) return fail(ctx, B200_ERR_INPUT_VALIDATION, "b200_kernel_scope_end without a scope");
	const int32_t rc = run_trace(ctx);
	ctx->tracing = false;
	ctx->trace.clear();
	ctx->mu.unlock();
	return rc;
}

int32_t b200_kernel_decl_value(b200_ctx *ctx, const uint64_t init[2], uint32_t *slot) {
	B200_LOCK_REC(ctx);
	B200_FLUSH(ctx);
	if (!ctx || !slot) return B200_ERR_INPUT_VALIDATION;
	int32_t rc = new_slot(ctx, slot);
	if (rc) return rc;
	if (ctx->tracing) {
		b200_trace_op op;
		op.kind = 0, op.slot = *slot, op.init[0] = init[0], op.init[1] = init[1];
		ctx->trace.push_back(op);
		return B200_OK;
	}
	return kernel_decl_now(ctx, init, *slot);
}
int32_t b200_kernel_sum_composition_evals(b200_ctx *ctx, const b200_dev_ptr *inputs, uint32_t n_inputs, uint64_t row_len, const b200_expr *expr, const uint64_t coeff[2], uint32_t slot) {
	B200_LOCK_REC(ctx);
	B200_FLUSH(ctx);
	if (!ctx || !expr) return B200_ERR_INPUT_VALIDATION;
	if (slot >= ctx->n_results) return fail(ctx, B200_ERR_INPUT_VALIDATION, "value slot %u not declared", slot);
	if (expr->n_vars > n_inputs) return fail(ctx, B200_ERR_INPUT_VALIDATION, "composition not match with input");
	if (ctx->tracing) {
		if (row_len == 0) return B200_OK;
		b200_trace_op op;
		op.kind = 1, op.slot = slot, op.row_len = row_len, op.expr = expr, op.coeff[0] = coeff[0], op.coeff[1] = coeff[1];
		op.inputs.assign(inputs, inputs + n_inputs);
		ctx->trace.push_back(std::move(op));
		return B200_OK;
	}
	return kernel_sum_now(ctx, inputs, n_inputs, row_len, expr, coeff, slot);
}
int32_t b200_kernel_add(b200_ctx *ctx, uint32_t log_len, b200_dev_ptr a, b200_dev_ptr b, b200_dev_ptr dst) {
	B200_LOCK_REC(ctx);
	B200_FLUSH(ctx);
	if (!ctx || log_len > 60) return B200_ERR_INPUT_VALIDATION;
	if (ctx->tracing) {
		b200_trace_op op;
		op.kind = 2, op.log_len = log_len, op.a = a, op.b = b, op.dst = dst;
		ctx->trace.push_back(op);
		return B200_OK;
	}
	return kernel_add_now(ctx, log_len, a, b, dst);
}
int32_t b200_kernel_add_assign(b200_ctx *ctx, uint32_t log_len, b200_dev_ptr src, b200_dev_ptr dst) {
	return b200_kernel_add(ctx, log_len, dst, src, dst);
}

int32_t b200_bivariate_round_evals(b200_ctx *ctx, const b200_dev_ptr *mls, uint32_t m, uint32_t n_vars, const uint32_t *ia, const uint32_t *ib, uint32_t n_comp, const uint64_t coeff[2], uint32_t *slot_y1, uint32_t *slot_yinf) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx || !slot_y1 || !slot_yinf) return B200_ERR_INPUT_VALIDATION;
	if (n_vars == 0 || n_vars > 60) return fail(ctx, B200_ERR_INPUT_VALIDATION, "n_vars must be in [1, 60]");
	for (uint32_t c = 0; c < n_comp; c++)
		if (ia[c] >= m || ib[c] >= m) return fail(ctx, B200_ERR_INPUT_VALIDATION, "composition %u indexes a missing multilinear", c);
	int32_t rc = new_slot(ctx, slot_y1);
	if (rc) return rc;
	rc = new_slot(ctx, slot_yinf);
	if (rc) return rc;
	if (n_comp == 0) return B200_OK;
	if (n_comp > 65535) return fail(ctx, B200_ERR_INPUT_VALIDATION, "too many compositions");
	uint64_t half = 1ull << (n_vars - 1);
	// batch coefficient powers alpha^c (host; n_comp scalar products)
	std::vector<uint4> pows(n_comp);
	hostf::u128 a = hostf::from_words(coeff), pw = 1;
	for (uint32_t c = 0; c < n_comp; c++) {
		uint64_t w[2] = {(uint64_t)pw, (uint64_t)(pw >> 64)};
		pows[c] = to_u4(w);
		pw = hostf::mul128(pw, a);
	}
	const int tc_mode = ctx->tune_round_evals_tc;
	if (tc_mode && half >= 4096 && half % tc::CHUNK == 0) {
		// tensor-core path (roundevals_tc.cuh): two inner-product jobs per composition
		std::vector<tc::TcJob> jobs(2 * (size_t)n_comp);
		std::vector<tc::TcTarget> targets(jobs.size());
		for (uint32_t c = 0; c < n_comp; c++) {
			const uint4 *a = (const uint4 *)mls[ia[c]], *b = (const uint4 *)mls[ib[c]];
			jobs[2 * c] = tc::TcJob{a + half, nullptr, b + half, nullptr};
			jobs[2 * c + 1] = tc::TcJob{a + half, a, b + half, b};
			targets[2 * c] = tc::TcTarget{2 * c, *slot_y1, pows[c]};
			targets[2 * c + 1] = tc::TcTarget{2 * c + 1, *slot_yinf, pows[c]};
		}
		return launch_tc_pairs(ctx, jobs, half, targets);
	}
	ArgPack pack;
	size_t o_m = pack.add(mls, sizeof(void *) * m), o_a = pack.add(ia, 4 * n_comp), o_b = pack.add(ib, 4 * n_comp), o_p = pack.add(pows.data(), 16 * n_comp);
	uint8_t *dbase;
	if ((rc = pack.commit(ctx, &dbase))) return rc;
	void *dm = dbase + o_m, *dia = dbase + o_a, *dib = dbase + o_b, *dp = dbase + o_p;
	uint32_t gx = grid_for(ctx, half, 256, 2);
	gx = std::max(1u, std::min(gx, (uint32_t)(ctx->n_sms * 4 / std::max(1u, std::min(n_comp, (uint32_t)ctx->n_sms * 4)) + 1)));
	k_bivariate_round_evals<<<dim3(gx, n_comp), 256, FIELD_TABLE_BYTES, ctx->stream>>>(ctx->d_tables, (const uint4 *const *)dm, half, (const uint32_t *)dia, (const uint32_t *)dib, (const uint4 *)dp, ctx->d_results + *slot_y1, ctx->d_results + *slot_yinf);
	B200_LAUNCH_CHECK(ctx);
	return B200_OK;
}

int32_t b200_eq_ind_round_evals(b200_ctx *ctx, const b200_dev_ptr *mls, const uint64_t *lens, const uint64_t *suffix_evals, uint32_t m, uint32_t n_vars, b200_dev_ptr eq_ind, const b200_expr *const *comps, const b200_expr *const *leads, uint32_t n_comp, const uint32_t *codes, const uint64_t *points, uint32_t n_points, uint32_t *first_slot) {
	if (ctx && !eq_ind) return fail(ctx, B200_ERR_INPUT_VALIDATION, "eq_ind_partial_evals is required by the eq-ind evaluator");
	return b200_sumcheck_round_evals(ctx, B200_HIGH_TO_LOW, mls, lens, suffix_evals, m, n_vars, eq_ind, comps, leads, n_comp, codes, points, n_points, first_slot);
}

int32_t b200_sumcheck_round_evals(b200_ctx *ctx, uint32_t order, const b200_dev_ptr *mls, const uint64_t *lens, const uint64_t *suffix_evals, uint32_t m, uint32_t n_vars, b200_dev_ptr eq_ind, const b200_expr *const *comps, const b200_expr *const *leads, uint32_t n_comp, const uint32_t *codes, const uint64_t *points, uint32_t n_points, uint32_t *first_slot) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx || !first_slot) return B200_ERR_INPUT_VALIDATION;
	if (n_vars == 0 || n_vars > 60) return fail(ctx, B200_ERR_INPUT_VALIDATION, "n_vars must be in [1, 60]");
	if (order != B200_LOW_TO_HIGH && order != B200_HIGH_TO_LOW) return fail(ctx, B200_ERR_INPUT_VALIDATION, "unknown evaluation order %u", order);
	if ((uint64_t)n_comp * n_points > 65535) return fail(ctx, B200_ERR_INPUT_VALIDATION, "too many (composition, point) pairs");
	for (uint32_t c = 0; c < n_comp; c++)
		if (!comps[c] || !leads[c] || comps[c]->n_vars > m || leads[c]->n_vars > m) return fail(ctx, B200_ERR_INPUT_VALIDATION, "composition %u does not match the multilinears", c);
	for (uint32_t p = 0; p < n_points; p++)
		if (codes[p] == 0) return fail(ctx, B200_ERR_INPUT_VALIDATION, "evaluation point code 0 (evaluate at 0) is never computed by the prover");
	uint32_t total = n_comp * n_points;
	if (ctx->n_results + total > MAX_RESULTS) return fail(ctx, B200_ERR_ALLOC, "out of result slots");
	*first_slot = ctx->n_results;
	ctx->n_results += total;
	if (total == 0) return B200_OK;
	std::vector<uint4> hp(n_points);
	for (uint32_t p = 0; p < n_points; p++) hp[p] = to_u4(points + 2 * p);
	EqIndArgs A;
	int32_t rc;
	std::vector<uint64_t> hlen(std::max(m, 1u));
	std::vector<uint4> hs(std::max(m, 1u));
	for (uint32_t t = 0; t < m; t++) {
		hlen[t] = lens ? lens[t] : (1ull << n_vars);
		if (hlen[t] > (1ull << n_vars)) return fail(ctx, B200_ERR_INPUT_VALIDATION, "multilinear %u: stored length exceeds 2^n_vars", t);
		hs[t] = suffix_evals ? to_u4(suffix_evals + 2 * t) : make_uint4(0, 0, 0, 0);
	}
	const int tc_mode = ctx->tune_round_evals_tc;
	const uint64_t half = 1ull << (n_vars - 1);
	if (tc_mode >= 1 && tc_mode != 2 && half >= 4096 && half % tc::CHUNK == 0 && order == B200_HIGH_TO_LOW) {
		// points 1 / infinity only, full-length multilinears, degree <= 2: monomial plan, no interpreter
		bool ok = true;
		for (uint32_t t = 0; t < m && ok; t++) ok = hlen[t] == 2 * half;
		for (uint32_t p = 0; p < n_points && ok; p++) {
			ok = codes[p] == 1 || codes[p] == 2;
			for (uint32_t c = 0; c < n_comp && ok; c++) ok = codes[p] == 1 ? comps[c]->poly_ok : leads[c]->poly_ok;
		}
		if (ok) return eq_ind_monomial_plan(ctx, mls, half, eq_ind, comps, leads, n_comp, codes, n_points, *first_slot);
	}
	// the interpreter kernels read the steps on the device (uploaded on first use)
	upload_exprs({{comps, n_comp}, {leads, n_comp}});
	std::vector<DevExpr> hc(n_comp), hl(n_comp);
	for (uint32_t c = 0; c < n_comp; c++) {
		hc[c] = dev_expr(comps[c]);
		hl[c] = dev_expr(leads[c]);
		B200_EXPR_READY(ctx, hc[c]);
		B200_EXPR_READY(ctx, hl[c]);
	}
	ArgPack pack;
	size_t o_m = pack.add(mls, sizeof(void *) * m), o_l = pack.add(hlen.data(), 8 * hlen.size()), o_s = pack.add(hs.data(), 16 * hs.size());
	size_t o_c = pack.add(hc.data(), sizeof(DevExpr) * n_comp), o_ld = pack.add(hl.data(), sizeof(DevExpr) * n_comp);
	size_t o_k = pack.add(codes, 4 * n_points), o_p = pack.add(hp.data(), 16 * n_points);
	uint8_t *dbase;
	if ((rc = pack.commit(ctx, &dbase))) return rc;
	A.mls = (const uint4 *const *)(dbase + o_m);
	A.lens = (const uint64_t *)(dbase + o_l);
	A.suffix = (const uint4 *)(dbase + o_s);
	A.n_mls = m;
	A.half = 1ull << (n_vars - 1);
	A.eq_ind = (const uint4 *)eq_ind;
	A.comps = (const DevExpr *)(dbase + o_c);
	A.comps_lead = (const DevExpr *)(dbase + o_ld);
	A.codes = (const uint32_t *)(dbase + o_k);
	A.points = (const uint4 *)(dbase + o_p);
	A.n_points = n_points;
	A.slots = ctx->d_results + *first_slot;
	A.low_to_high = order == B200_LOW_TO_HIGH;
	uint32_t gx = grid_for(ctx, A.half, 256, 2);
	gx = std::max(1u, std::min(gx, (uint32_t)(ctx->n_sms * 4 / std::min(total, (uint32_t)ctx->n_sms * 4) + 1)));
	const uint64_t vals_bytes = (uint64_t)total * A.half * 16;
	// materialise C(P(i)), then sum_i E[i] * val[i] as tensor-core inner-product jobs -- when the values fit: at most
	// half of the free device memory beyond the scratch already held, and a failed allocation falls through to the
	// scratch-free per-lane kernel instead of failing the round
	const uint64_t gmat_bytes = (uint64_t)(total + 1) * 512 * 4;
	bool materialise = tc_mode && eq_ind && A.half >= 4096 && A.half % tc::CHUNK == 0 && vals_bytes <= (48ull << 30);
	if (materialise && vals_bytes + gmat_bytes > ctx->scratch_bytes) {
		size_t free_b = 0, total_b = 0;
		if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || vals_bytes + gmat_bytes > (free_b + ctx->scratch_bytes) / 2) materialise = false;
		cudaGetLastError();
	}
	if (materialise && ensure_scratch(ctx, vals_bytes + gmat_bytes) != B200_OK) materialise = false;
	if (materialise) {
		uint4 *vals = (uint4 *)ctx->d_scratch;
		k_eq_ind_vals<<<dim3(gx, total), 256, FIELD_TABLE_BYTES, ctx->stream>>>(ctx->d_tables, A, vals);
		B200_LAUNCH_CHECK(ctx);
		std::vector<tc::TcJob> jobs(total);
		std::vector<tc::TcTarget> targets(total);
		for (uint32_t j = 0; j < total; j++) {
			jobs[j] = tc::TcJob{(const uint4 *)eq_ind, nullptr, vals + (uint64_t)j * A.half, nullptr};
			targets[j] = tc::TcTarget{j, *first_slot + j, make_uint4(1, 0, 0, 0)};
		}
		return launch_tc_pairs(ctx, jobs, A.half, targets, vals_bytes);
	}
	k_eq_ind_round_evals<<<dim3(gx, total), 256, FIELD_TABLE_BYTES, ctx->stream>>>(ctx->d_tables, A);
	B200_LAUNCH_CHECK(ctx);
	return B200_OK;
}

// ---- persistent eq-ind sumcheck (tail_grid.cuh) ----------------------------------------------------------
int32_t b200_sumcheck_tail_start(b200_ctx *ctx, const b200_dev_ptr *mls, uint32_t m, uint32_t n_vars, b200_dev_ptr eq_ind, const b200_expr *const *comps,
								 const b200_expr *const *leads, uint32_t n_comp, const uint32_t *codes, const uint64_t *points, uint32_t n_points,
								 uint32_t first_round_skip, b200_tail **out) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx || !out || !eq_ind) return B200_ERR_INPUT_VALIDATION;
	if (n_vars == 0 || n_vars > 28) return fail(ctx, B200_ERR_INPUT_VALIDATION, "sumcheck tail: n_vars must be in [1, 28]");
	if (first_round_skip >= n_points) return fail(ctx, B200_ERR_INPUT_VALIDATION, "sumcheck tail: the first round must keep at least one evaluation point");
	const uint32_t n_vals = n_comp * n_points;
	if (n_vals == 0 || n_vals > TG_MAX_VALS) return fail(ctx, B200_ERR_INPUT_VALIDATION, "sumcheck tail: 1..4096 (composition, point) pairs");
	for (uint32_t c = 0; c < n_comp; c++)
		if (!comps[c] || !leads[c] || comps[c]->n_vars > m || leads[c]->n_vars > m) return fail(ctx, B200_ERR_INPUT_VALIDATION, "composition %u does not match the multilinears", c);
	for (uint32_t p = 0; p < n_points; p++)
		if (codes[p] == 0) return fail(ctx, B200_ERR_INPUT_VALIDATION, "evaluation point code 0 is never computed by the prover");
	if (ctx->tail_active) return fail(ctx, B200_ERR_INPUT_VALIDATION, "a sumcheck tail is already running on this context");
	std::unique_ptr<b200_tail> t(new b200_tail);
	t->ctx = ctx, t->n_vars = n_vars, t->n_vals = n_vals;
	t->off_chal = (size_t)n_vars * n_vals * 32;
	t->off_status = t->off_chal + 32 * (size_t)n_vars;
	t->off_trace = t->off_status + 32;
	const size_t bytes = t->off_trace + 32 * (size_t)n_vars;
	// one host-mapped mailbox per context, allocated on first use (pinned allocations cost ~100 us: not per sumcheck)
	if (ctx->tail_mb_bytes < bytes) {
		if (ctx->h_tail_mb) {
			B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
			cudaFreeHost(ctx->h_tail_mb);
			ctx->h_tail_mb = nullptr, ctx->tail_mb_bytes = 0;
		}
		const size_t want = std::max<size_t>(bytes, 256u << 10);
		if (cudaHostAlloc((void **)&ctx->h_tail_mb, want, cudaHostAllocMapped) != cudaSuccess) {
			cudaGetLastError();
			return fail(ctx, B200_ERR_ALLOC, "out of pinned host memory (sumcheck tail mailbox)");
		}
		if (cudaHostGetDevicePointer((void **)&ctx->d_tail_mb, ctx->h_tail_mb, 0) != cudaSuccess) {
			cudaGetLastError();
			cudaFreeHost(ctx->h_tail_mb);
			ctx->h_tail_mb = nullptr;
			return fail(ctx, B200_ERR_DEVICE, "host-mapped memory is not available");
		}
		ctx->tail_mb_bytes = want;
	}
	t->h_mb = ctx->h_tail_mb, t->d_mb = ctx->d_tail_mb;
	memset(t->h_mb, 0, bytes);  // every slot is validated by its complement: zero = not there yet
	upload_exprs({{comps, n_comp}, {leads, n_comp}});
	std::vector<DevExpr> hc(n_comp), hl(n_comp);
	std::vector<uint32_t> step_off(2 * n_comp + 1, 0);
	for (uint32_t c = 0; c < n_comp; c++) {
		hc[c] = dev_expr(comps[c]);
		hl[c] = dev_expr(leads[c]);
		if ((hc[c].n_steps && !hc[c].steps) || (hl[c].n_steps && !hl[c].steps)) return fail(ctx, B200_ERR_ALLOC, "out of device memory (expression steps)");
	}
	for (uint32_t e = 0; e < 2 * n_comp; e++) step_off[e + 1] = step_off[e] + (e < n_comp ? hc[e] : hl[e - n_comp]).n_steps;
	const bool cache_steps = 32 * (size_t)step_off[2 * n_comp] + sizeof(DevExpr) * 2 * n_comp <= TG_STEP_CACHE_BYTES;
	std::vector<uint4> hp(n_points);
	for (uint32_t p = 0; p < n_points; p++) hp[p] = to_u4(points + 2 * p);
	ArgPack pack;
	size_t o_m = pack.add(mls, sizeof(void *) * m), o_c = pack.add(hc.data(), sizeof(DevExpr) * n_comp), o_l = pack.add(hl.data(), sizeof(DevExpr) * n_comp);
	size_t o_k = pack.add(codes, 4 * n_points), o_p = pack.add(hp.data(), 16 * n_points), o_s = pack.add(step_off.data(), 4 * step_off.size());
	uint8_t *dbase;
	int32_t rc = pack.commit(ctx, &dbase);
	if (rc) return rc;
	TailGridArgs GA;
	TailArgs &A = GA.t;
	A.mls = (uint4 *const *)(dbase + o_m), A.m = m, A.n_vars = n_vars, A.eq_ind = (uint4 *)eq_ind;
	A.comps = (const DevExpr *)(dbase + o_c), A.leads = (const DevExpr *)(dbase + o_l), A.n_comp = n_comp, A.n_points = n_points;
	A.codes = (const uint32_t *)(dbase + o_k), A.points = (const uint4 *)(dbase + o_p);
	A.mb_vals = (uint4 *)t->d_mb, A.mb_chal = (uint4 *)(t->d_mb + t->off_chal);
	A.status = (volatile uint32_t *)(t->d_mb + t->off_status);
	A.mb_trace = ctx->tune_tail_trace ? (uint64_t *)(t->d_mb + t->off_trace) : nullptr;
	A.timeout_ns = 5ull * 1000 * 1000 * 1000;
	// hypercube chunks of 32 indices over a co-resident grid (one CTA when there is a single chunk, or on request)
	const uint64_t n_chunks = ((1ull << (n_vars - 1)) + 31) >> 5;
	uint32_t G = 1;
	if (ctx->tune_tail_grid)
		while (G * 2 <= std::min<uint64_t>({(uint64_t)TG_MAX_CTAS, n_chunks, (uint64_t)ctx->n_sms})) G *= 2;
	constexpr size_t WS_ACC = 28 * (size_t)TG_MAX_VALS * 16, WS_BYTES = WS_ACC + 28 * 16 + 28 * 4 + 16;
	if (!ctx->d_tail_ws) B200_CUDA(ctx, cudaMalloc((void **)&ctx->d_tail_ws, WS_BYTES));
	if (G > 1) B200_CUDA(ctx, cudaMemsetAsync(ctx->d_tail_ws, 0, (size_t)n_vars * n_vals * 16, ctx->stream));
	B200_CUDA(ctx, cudaMemsetAsync(ctx->d_tail_ws + WS_ACC, 0, WS_BYTES - WS_ACC, ctx->stream));
	GA.acc = (uint4 *)ctx->d_tail_ws, GA.g_chal = (uint4 *)(ctx->d_tail_ws + WS_ACC);
	GA.g_chal_seq = (uint32_t *)(ctx->d_tail_ws + WS_ACC + 28 * 16), GA.bar = GA.g_chal_seq + 28, GA.g_abort = GA.bar + 1;
	GA.first_skip = first_round_skip;
	GA.step_off = cache_steps ? (const uint32_t *)(dbase + o_s) : nullptr;
	GA.off_acc = (FIELD_TABLE_BYTES + 127) & ~127u;
	GA.off_ex = GA.off_acc + 16 * n_vals;
	GA.off_steps = GA.off_ex + (cache_steps ? (uint32_t)sizeof(DevExpr) * 2 * n_comp : 0);
	size_t smem_bytes = GA.off_steps + (cache_steps ? 32 * (size_t)step_off[2 * n_comp] : 0);
	GA.off_hdr = GA.off_k64 = 0;
	if (m <= 2048 && n_points <= 64) {
		GA.off_hdr = (uint32_t)smem_bytes;
		smem_bytes += 16 * (size_t)n_points + 8 * (size_t)m + 4 * (size_t)n_points;
		smem_bytes = (smem_bytes + 127) & ~(size_t)127;
	}
	GA.res_half = 0;
	if (smem_bytes + LUT_BYTES + 1536 <= TG_SMEM_MAX) {
		GA.off_k64 = (uint32_t)smem_bytes;
		smem_bytes += LUT_BYTES + 1536;
		// the small rounds run on CTA 0 from shared-memory copies held in the same region: (2 m + 1) * half elements
		if (GA.off_hdr && ctx->tune_tail_resident)
			// (128 pairs: with 256 one SM alone is slower than eight SMs plus two barriers -- 17.6 us against 11.6 for config #3)
			for (uint32_t h = 128; h >= 1; h >>= 1)
				if ((2 * (uint64_t)m + 1) * h * 16 <= LUT_BYTES + 1536) {
					GA.res_half = h;
					break;
				}
	}
	const uint8_t *tables = ctx->d_tables;
	void *kargs[] = {(void *)&tables, (void *)&GA};
	B200_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)k_sumcheck_tail_grid, dim3(G), dim3(TG_THREADS), kargs, smem_bytes, ctx->stream));
	B200_LAUNCH_CHECK(ctx);
	ctx->tail_active = true;
	*out = t.release();
	return B200_OK;
}
// the values of the next round (n_comp * n_points elements, [composition][point]); blocks until the kernel posts them
int32_t b200_sumcheck_tail_round_evals(b200_tail *t, uint64_t *out) {
	if (!t || !out) return B200_ERR_INPUT_VALIDATION;
	b200_ctx *ctx = t->ctx;
	if (t->round_out >= t->n_vars || t->round_out != t->round_in) return fail(ctx, B200_ERR_INPUT_VALIDATION, "sumcheck tail: round values requested out of order");
	const uint32_t r = t->round_out;
	const volatile uint64_t *slot = (const volatile uint64_t *)(t->h_mb + (size_t)r * t->n_vals * 32);
	volatile uint32_t *status = (volatile uint32_t *)(t->h_mb + t->off_status);
	const auto t0 = std::chrono::steady_clock::now();
	for (uint32_t v = 0; v < t->n_vals; v++) {
		uint64_t lo, hi;
		for (uint64_t spins = 0;; spins++) {
			lo = slot[4 * v], hi = slot[4 * v + 1];
			if (lo == ~slot[4 * v + 2] && hi == ~slot[4 * v + 3]) break;
			if (*status) return fail(ctx, B200_ERR_DEVICE, "sumcheck tail: the kernel's watchdog expired");
			if ((spins & 0xFFFF) == 0xFFFF) {
				if (cudaStreamQuery(ctx->stream) != cudaErrorNotReady && !(slot[4 * v] == ~slot[4 * v + 2] && slot[4 * v + 1] == ~slot[4 * v + 3]))
					return fail(ctx, B200_ERR_DEVICE, "sumcheck tail: the kernel ended before posting round %u", r);
				if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(10)) return fail(ctx, B200_ERR_DEVICE, "sumcheck tail: timed out waiting for round %u", r);
			}
		}
		out[2 * v] = lo, out[2 * v + 1] = hi;
	}
	t->round_out++;
	return B200_OK;
}
// post the challenge of the round whose values were just read: the kernel folds and goes on to the next round
int32_t b200_sumcheck_tail_challenge(b200_tail *t, const uint64_t z[2]) {
	if (!t || !z) return B200_ERR_INPUT_VALIDATION;
	if (t->round_in + 1 != t->round_out) return fail(t->ctx, B200_ERR_INPUT_VALIDATION, "sumcheck tail: challenge posted out of order");
	volatile uint64_t *slot = (volatile uint64_t *)(t->h_mb + t->off_chal + 32 * (size_t)t->round_in);
	slot[2] = ~z[0], slot[3] = ~z[1];
	slot[0] = z[0], slot[1] = z[1];
	std::atomic_thread_fence(std::memory_order_release);
	t->round_in++;
	return B200_OK;
}
// wait for the kernel to end (all challenges posted) and release the mailbox; the multilinears hold their final folds
int32_t b200_sumcheck_tail_finish(b200_tail *t) {
	if (!t) return B200_ERR_INPUT_VALIDATION;
	b200_ctx *ctx = t->ctx;
	B200_LOCK(ctx);
	int32_t rc = B200_OK;
	if (t->round_in != t->n_vars) {
		// abandoned: the watchdog ends the kernel
		rc = fail(ctx, B200_ERR_INPUT_VALIDATION, "sumcheck tail finished after %u of %u challenges", t->round_in, t->n_vars);
	}
	cudaError_t e = cudaStreamSynchronize(ctx->stream);
	if (e != cudaSuccess && rc == B200_OK) rc = fail(ctx, B200_ERR_DEVICE, "sumcheck tail: %s", cudaGetErrorString(e));
	if (rc == B200_OK && *(volatile uint32_t *)(t->h_mb + t->off_status)) rc = fail(ctx, B200_ERR_DEVICE, "sumcheck tail: the kernel's watchdog expired");
	if (ctx->tune_tail_trace) {
		// debugging aid (b200_ctx_set_tuning "tail_trace"): CTA 0's %globaltimer stamps per round
		const uint64_t *tr = (const uint64_t *)(t->h_mb + t->off_trace);
		for (uint32_t r = 0; r < t->n_vars; r++)
			fprintf(stderr, "tail round %2u: values %6.2f us, mailbox %6.2f us, fold %6.2f us%s\n", r, (tr[4 * r + 1] - tr[4 * r]) * 1e-3, (tr[4 * r + 2] - tr[4 * r + 1]) * 1e-3,
					r + 1 < t->n_vars ? (tr[4 * r + 3] - tr[4 * r + 2]) * 1e-3 : 0.0, r + 1 < t->n_vars ? "" : " (last)");
	}
	ctx->tail_active = false;
	for (void *p : ctx->deferred_free) cudaFree(p);
	ctx->deferred_free.clear();
	delete t;
	return rc;
}

// ---- NTT -----------------------------------------------------------------------------------------
int32_t b200_ntt_create(b200_ctx *ctx, uint32_t kt, uint32_t d, b200_ntt **out) {
	B200_LOCK(ctx);
	if (!ctx || !out) return B200_ERR_INPUT_VALIDATION;
	if (kt < 3 || kt > 5) return fail(ctx, B200_ERR_INPUT_VALIDATION, "twiddle field must be B8, B16 or B32");
	if (d == 0) return fail(ctx, B200_ERR_NTT_DOMAIN, "domain size is less than 2**1");
	if (d > (1u << kt)) return fail(ctx, B200_ERR_NTT_FIELD, "field order must be at least 2**%u", d);
	std::unique_ptr<b200_ntt> n(new b200_ntt);
	n->ctx = ctx;
	n->kt = kt;
	n->d = d;
	// precompute_subspace_evals (crates/ntt/src/twiddle.rs:244-313) over the basis beta_j = 1 << j
	using hostf::u128;
	std::vector<std::vector<u128>> s(d);
	std::vector<u128> norm(d);
	norm[0] = 1;
	for (uint32_t j = 1; j < d; j++) s[0].push_back((u128)1 << j);
	for (uint32_t r = 1; r < d; r++) {
		u128 np = norm[r - 1];
		auto phi = [&](u128 e) { return hostf::mul(e, e, kt) ^ hostf::mul(np, e, kt); };
		norm[r] = phi(s[r - 1][0]);
		for (size_t j = 1; j < s[r - 1].size(); j++) s[r].push_back(phi(s[r - 1][j]));
	}
	std::vector<uint32_t> flat(32 * 32, 0);
	n->s_evals.resize(d);
	for (uint32_t r = 0; r < d; r++) {
		u128 inv = hostf::invert(norm[r], kt);
		for (size_t j = 0; j < s[r].size(); j++) {
			u128 v = hostf::mul(s[r][j], inv, kt);
			n->s_evals[r].push_back((uint64_t)v);
			flat[r * 32 + j] = (uint32_t)v;
		}
	}
	if (cudaMalloc(&n->d_s_evals, sizeof(uint32_t) * 32 * 32) != cudaSuccess) {
		cudaGetLastError();
		return fail(ctx, B200_ERR_ALLOC, "out of device memory");
	}
	B200_CUDA(ctx, cudaMemcpy(n->d_s_evals, flat.data(), sizeof(uint32_t) * 32 * 32, cudaMemcpyHostToDevice));
	if (kt == 5) {
		// basis products of the look-up-table passes (ntt_lut.cuh): every table entry is an XOR of these
		std::vector<uint32_t> basis(32 * 32 * 32, 0);
		for (uint32_t r = 0; r < d; r++)
			for (size_t j = 0; j < n->s_evals[r].size(); j++)
				for (uint32_t k = 0; k < 32; k++) basis[(r * 32 + j) * 32 + k] = (uint32_t)hostf::mul((u128)n->s_evals[r][j], (u128)1 << k, 5);
		if (cudaMalloc(&n->d_basis, sizeof(uint32_t) * basis.size()) != cudaSuccess) {
			cudaGetLastError();
			cudaFree(n->d_s_evals);
			return fail(ctx, B200_ERR_ALLOC, "out of device memory");
		}
		B200_CUDA(ctx, cudaMemcpy(n->d_basis, basis.data(), sizeof(uint32_t) * basis.size(), cudaMemcpyHostToDevice));
	}
	*out = n.release();
	return B200_OK;
}
void b200_ntt_destroy(b200_ntt *ntt) {
	if (!ntt) return;
	ctx_free(ntt->ctx, ntt->d_s_evals);
	ctx_free(ntt->ctx, ntt->d_basis);
	delete ntt;
}
uint32_t b200_ntt_log_domain_size(const b200_ntt *ntt) { return ntt ? ntt->d : 0; }

int32_t b200_ntt_get_subspace_eval(const b200_ntt *ntt, uint32_t i, uint64_t j, uint64_t out[2]) {
	if (!ntt || !out || i == 0 || i > ntt->d) return B200_ERR_INPUT_VALIDATION;
	const auto &row = ntt->s_evals[ntt->d - i];
	uint64_t t = 0;
	for (size_t b = 0; b < row.size(); b++)
		if ((j >> b) & 1) t ^= row[b];
	out[0] = t;
	out[1] = 0;
	return B200_OK;
}

static int32_t ntt_run(b200_ctx *ctx, const b200_ntt *ntt, int inverse, b200_dev_ptr data, uint32_t kd, uint64_t n_elems, uint32_t log_x, uint32_t log_y, uint32_t log_z, uint64_t coset, uint32_t coset_bits, uint32_t skip_rounds) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx || !ntt) return B200_ERR_INPUT_VALIDATION;
	if (kd < ntt->kt || kd > 7) return fail(ctx, B200_ERR_INPUT_VALIDATION, "element width 2^%u bits is not an extension of the twiddle field", kd);
	// check_batch_transform_inputs_and_params (crates/ntt/src/single_threaded.rs:364-406), WIDTH = 1
	if (!is_pow2(n_elems)) return fail(ctx, B200_ERR_NTT_POWER_OF_TWO, "the input length must be a power of two");
	if (skip_rounds > log_y) return fail(ctx, B200_ERR_NTT_SKIP_ROUNDS, "the skip_rounds parameter exceeds the total number of NTT rounds");
	if (log_x + log_y + log_z > 62) return fail(ctx, B200_ERR_NTT_BATCH, "the batch size is greater than the number of elements");
	uint64_t full_y = n_elems >> (log_x + log_z);
	if (((1ull << log_y) != full_y && n_elems > 2) || (1ull << log_y) > full_y) return fail(ctx, B200_ERR_NTT_BATCH, "the batch size is greater than the number of elements");
	if (coset_bits > 62 || coset >= (1ull << coset_bits)) return fail(ctx, B200_ERR_NTT_COSET, "coset index must be less than 2**%u, got %llu", coset_bits, (unsigned long long)coset);
	if (log_y + coset_bits > ntt->d) return fail(ctx, B200_ERR_NTT_DOMAIN, "domain size is less than 2**%u", log_y + coset_bits);
	uint32_t n_layers = log_y - skip_rounds;
	if (n_layers == 0) return B200_OK;
	uint32_t lx = log_x + (kd - ntt->kt);  // *_transform_ext: extension limbs become extra batch columns
	// pass plan, bottom-up.  kind 0 = table-driven scalar pass (ntt.cuh), kind 1 = bit-sliced pass
	// (ntt_bs.cuh; needs >= 32 contiguous positions sharing every twiddle: lx + i_lo >= 5, B32 only).
	struct Pass { uint32_t kind, i_lo, R, R_exec, log_c; };
	// bit-sliced tile size (ntt_bs.cuh): 512-unit tiles up to 2^25 coefficients, 256-unit tiles above
	const uint32_t LT = ((n_elems << (kd - ntt->kt)) >> 25) ? 8u : NTT_BS_MAX_LOG_TILE;
	std::vector<Pass> plan;
	const uint32_t MAX_LOG_TILE = 13;
	// ctx->tune_ntt: 0 = look-up-table passes + bit-sliced low layers (default), 1 = bit-sliced only, 2 = scalar tables
	if (ntt->kt == 5 && ctx->tune_ntt != 2) {
		// Look-up-table passes (ntt_lut.cuh) take every layer whose twiddles are shared by >= 2^5 positions
		// (lx + i >= 5; 2^6 when the bit-sliced low pass exists anyway); the layers below stay bit-sliced.
		uint32_t n_bs = lx >= 5 ? 0u : std::min(n_layers, 6u - lx);
		uint32_t n_lut = n_layers - n_bs;
		if (ctx->tune_ntt == 1 || n_lut < 3 || !ntt->d_basis) n_bs = n_layers, n_lut = 0;
		uint32_t i_lo = 0;
		if (n_bs && lx < 5 && lx + log_y >= 5) {
			// kind 2: lowest pass on units of 32 consecutive scalars (intra-unit layers + up to Rt inter-unit)
			uint32_t L0 = 5 - lx;
			uint32_t Rt = std::min(LT, lx + log_y - 5);
			uint32_t n_intra = std::min(n_bs, L0);
			uint32_t n_inter = std::min(n_bs - n_intra, Rt);
			uint32_t rest = n_bs - n_intra - n_inter;
			if (rest > 0 && rest < 4 && n_inter + rest >= 8) n_inter = n_inter + rest - 4;  // avoid a tiny upper pass
			plan.push_back(Pass{2, 0, Rt, n_intra, n_inter});
			i_lo = n_intra + n_inter;
		} else if (n_bs && lx < 5) {
			plan.push_back(Pass{0, 0, log_y, n_bs, lx});  // transform smaller than one unit
			i_lo = n_bs;
		}
		// tiles of 2^(R + log_cu) = 2^LT units: one butterfly-unit per thread and layer
		while (i_lo < n_bs) {
			const uint32_t rem = n_bs - i_lo;
			uint32_t log_cu = rem <= LT ? 0u : std::min(lx + i_lo - 5, 1u);  // a last pass of exactly LT layers takes single-unit rows
			uint32_t R = std::min(rem, LT - log_cu);
			if (n_bs - i_lo - R > 0 && n_bs - i_lo - R < 4) R = (n_bs - i_lo + 1) / 2;  // avoid a tiny last pass
			log_cu = std::min(lx + i_lo - 5, LT - R);  // short passes take wider tiles: always 2^LT units
			plan.push_back(Pass{1, i_lo, R, R, log_cu});
			i_lo += R;
		}
		if (n_lut) {
			// passes of 3..6 layers, the larger ones at the bottom (their table builds amortise over fewer columns)
			const uint32_t n_pass = (n_lut + 5) / 6, base = n_lut / n_pass, rem = n_lut % n_pass;
			for (uint32_t p = 0; p < n_pass; p++) {
				const uint32_t R = base + (p < rem ? 1u : 0u);
				plan.push_back(Pass{3, i_lo, R, R, std::min(lx + i_lo, ctx->tune_ntt_log_cc)});
				i_lo += R;
			}
		}
	} else {
		for (uint32_t i_lo = 0; i_lo < n_layers;) {
			uint32_t log_inner = lx + i_lo;
			uint32_t log_c = std::min(5u, log_inner);
			uint32_t R = std::min(n_layers - i_lo, MAX_LOG_TILE - log_c);
			if (log_c == 5) R = std::min(R, 8u);
			plan.push_back(Pass{0, i_lo, R, R, log_c});
			i_lo += R;
		}
	}
	uint32_t n_z = 1u << log_z;
	if (n_z > 65535) return fail(ctx, B200_ERR_INPUT_VALIDATION, "log_z too large");
	for (size_t pi = 0; pi < plan.size(); pi++) {
		const Pass &P = inverse ? plan[pi] : plan[plan.size() - 1 - pi];
		// neighbours in execution order: bit-sliced passes hand the data over bit-sliced
		auto sliced_at = [&](size_t q) {
			const uint32_t k = (inverse ? plan[q] : plan[plan.size() - 1 - q]).kind;
			return k == 1 || k == 2;
		};
		const uint32_t in_sliced = pi > 0 && sliced_at(pi) && sliced_at(pi - 1);
		const uint32_t out_sliced = pi + 1 < plan.size() && sliced_at(pi) && sliced_at(pi + 1);
		uint32_t row0 = ntt->d - (log_y + coset_bits);
		if (P.kind == 3) {
			nttl::Args L;
			L.data = (uint32_t *)data;
			L.basis = ntt->d_basis;
			L.lx = lx;
			L.log_y = log_y;
			L.i_lo = P.i_lo;
			L.row0 = row0;
			L.d = ntt->d;
			L.log_cc = P.log_c;
			L.n_z = n_z;
			L.nbits = std::min(32u, log_y - 1 - P.i_lo + coset_bits);
			L.coset = coset;
			const uint32_t R1 = P.R - 3;
			const uint64_t items = (uint64_t)n_z << (log_y - P.i_lo - P.R + lx + P.i_lo - P.log_c);
			// tables that live for many tiles: one CTA per SM, 16 compute warps, a deep tile ring; tables rebuilt
			// for every tile (a CTA-wide rendezvous): two CTAs of 8 compute warps per SM
			const bool per_tile_tables = lx + P.i_lo == P.log_c;
			L.cw = per_tile_tables || ctx->tune_ntt_cw == 8 ? 8 : nttl::MAX_CW;
			// byte tables (64 KiB more) for the two widest-shared layers when the tables live for many tiles
			const bool byte_tabs = !per_tile_tables && L.cw == nttl::MAX_CW && R1 >= 2 && ctx->tune_ntt_byte;
			L.nbuf = nttl::MAX_NBUF;
			const uint32_t budget = L.cw == 8 ? 112u * 1024u : 226u * 1024u;
			while (L.nbuf > 2 && nttl::layout((int)R1, P.log_c, L.nbits, L.nbuf, byte_tabs).total > budget) L.nbuf--;
			const uint32_t grid = (uint32_t)std::min<uint64_t>(items, (L.cw == 8 ? 2ull : 1ull) * ctx->n_sms);
			const uint32_t smem = nttl::layout((int)R1, P.log_c, L.nbits, L.nbuf, byte_tabs).total;
			const uint32_t threads = 32u * (L.cw + 2);
#define B200_NTT_LUT(R1V, BYTE)                                                                         \
	case R1V:                                                                                           \
		if (inverse) nttl::k_ntt_lut<R1V, true, BYTE><<<grid, threads, smem, ctx->stream>>>(L);         \
		else nttl::k_ntt_lut<R1V, false, BYTE><<<grid, threads, smem, ctx->stream>>>(L);                \
		break;
			if (byte_tabs) switch (R1) {
					B200_NTT_LUT(2, true)
					B200_NTT_LUT(3, true)
				}
			else switch (R1) {
					B200_NTT_LUT(0, false)
					B200_NTT_LUT(1, false)
					B200_NTT_LUT(2, false)
					B200_NTT_LUT(3, false)
				}
#undef B200_NTT_LUT
			B200_LAUNCH_CHECK(ctx);
			continue;
		}
		if (P.kind == 2) {
			NttBsLowArgs L;
			L.data = (uint32_t *)data;
			L.log_x = lx;
			L.log_y = log_y;
			L.Rt = P.R;
			L.n_intra = P.R_exec;
			L.n_inter = P.log_c;
			L.row0 = row0;
			L.d = ntt->d;
			L.coset = coset;
			L.inverse = inverse;
			L.s_evals = ntt->d_s_evals;
			L.in_sliced = in_sliced;
			L.out_sliced = out_sliced;
			uint64_t n_blocks = 1ull << (lx + log_y - 5 - P.R);
			if (n_blocks > 0x7fffffffull) return fail(ctx, B200_ERR_INPUT_VALIDATION, "transform too large");
			uint32_t smem = (184u << P.R) + 640 + 16;
			k_ntt_bs_low<<<dim3((uint32_t)n_blocks, n_z), ntt_bs_threads(LT), smem, ctx->stream>>>(L);
			B200_LAUNCH_CHECK(ctx);
			continue;
		}
		if (P.kind == 1) {
			NttBsArgs B;
			B.data = (uint32_t *)data;
			B.log_x = lx;
			B.log_y = log_y;
			B.i_lo = P.i_lo;
			B.R = P.R;
			B.log_cu = P.log_c;
			B.row0 = row0;
			B.d = ntt->d;
			B.coset = coset;
			B.inverse = inverse;
			B.s_evals = ntt->d_s_evals;
			B.in_sliced = in_sliced;
			B.out_sliced = out_sliced;
			uint64_t n_blocks = 1ull << (log_y - (P.i_lo + P.R) + (lx + P.i_lo - 5 - P.log_c));
			if (n_blocks > 0x7fffffffull) return fail(ctx, B200_ERR_INPUT_VALIDATION, "transform too large");
			uint32_t smem = (36u << P.R) + 16 + (128u << (P.R + P.log_c));
			k_ntt_bs_pass<<<dim3((uint32_t)n_blocks, n_z), ntt_bs_threads(LT), smem, ctx->stream>>>(B);
			B200_LAUNCH_CHECK(ctx);
			continue;
		}
		NttPassArgs A;
		A.data = data;
		A.log_x = lx;
		A.log_y = log_y;
		A.i_lo = P.i_lo;
		A.R = P.R;
		A.R_exec = P.R_exec;
		A.log_c = P.log_c;
		A.row0 = row0;
		A.d = ntt->d;
		A.coset = coset;
		A.inverse = inverse;
		A.s_evals = ntt->d_s_evals;
		uint64_t n_blocks = 1ull << (log_y - (P.i_lo + P.R) + (lx + P.i_lo - P.log_c));
		if (n_blocks > 0x7fffffffull) return fail(ctx, B200_ERR_INPUT_VALIDATION, "transform too large");
		uint32_t esz = 1u << (ntt->kt - 3);
		uint32_t smem = FIELD_TABLE_BYTES + (4u << P.R) + (esz << (P.R + P.log_c));
		dim3 grid((uint32_t)n_blocks, n_z);
		if (ntt->kt == 5) k_ntt_pass<uint32_t><<<grid, 256, smem, ctx->stream>>>(ctx->d_tables, A);
		else if (ntt->kt == 4) k_ntt_pass<uint16_t><<<grid, 256, smem, ctx->stream>>>(ctx->d_tables, A);
		else k_ntt_pass<uint8_t><<<grid, 256, smem, ctx->stream>>>(ctx->d_tables, A);
		B200_LAUNCH_CHECK(ctx);
	}
	return B200_OK;
}

int32_t b200_ntt_forward(b200_ctx *ctx, const b200_ntt *ntt, b200_dev_ptr data, uint32_t kd, uint64_t n, uint32_t lx, uint32_t ly, uint32_t lz, uint64_t coset, uint32_t cb, uint32_t skip) {
	return ntt_run(ctx, ntt, 0, data, kd, n, lx, ly, lz, coset, cb, skip);
}
int32_t b200_ntt_inverse(b200_ctx *ctx, const b200_ntt *ntt, b200_dev_ptr data, uint32_t kd, uint64_t n, uint32_t lx, uint32_t ly, uint32_t lz, uint64_t coset, uint32_t cb, uint32_t skip) {
	return ntt_run(ctx, ntt, 1, data, kd, n, lx, ly, lz, coset, cb, skip);
}

static int32_t ntt_host(b200_ctx *ctx, const b200_ntt *ntt, int inverse, void *host, uint32_t kd, uint64_t n, uint32_t lx, uint32_t ly, uint32_t lz, uint64_t coset, uint32_t cb, uint32_t skip) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx || !ntt) return B200_ERR_INPUT_VALIDATION;
	if (kd < 3 || kd > 7) return fail(ctx, B200_ERR_INPUT_VALIDATION, "bad element width");
	uint64_t bytes = n << (kd - 3);
	int32_t rc = ensure_scratch(ctx, std::max<uint64_t>(bytes, 16));
	if (rc) return rc;
	B200_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
	rc = ntt_run(ctx, ntt, inverse, ctx->d_scratch, kd, n, lx, ly, lz, coset, cb, skip);
	if (rc) return rc;
	B200_CUDA(ctx, cudaMemcpyAsync(host, ctx->d_scratch, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return B200_OK;
}
int32_t b200_ntt_forward_host(b200_ctx *ctx, const b200_ntt *ntt, void *host, uint32_t kd, uint64_t n, uint32_t lx, uint32_t ly, uint32_t lz, uint64_t coset, uint32_t cb, uint32_t skip) {
	return ntt_host(ctx, ntt, 0, host, kd, n, lx, ly, lz, coset, cb, skip);
}
int32_t b200_ntt_inverse_host(b200_ctx *ctx, const b200_ntt *ntt, void *host, uint32_t kd, uint64_t n, uint32_t lx, uint32_t ly, uint32_t lz, uint64_t coset, uint32_t cb, uint32_t skip) {
	return ntt_host(ctx, ntt, 1, host, kd, n, lx, ly, lz, coset, cb, skip);
}

int32_t b200_fri_fold(b200_ctx *ctx, const b200_ntt *ntt, uint32_t log_len, uint32_t log_batch, const uint64_t *challenges, uint32_t n_ch, b200_dev_ptr in, uint64_t n_in, b200_dev_ptr out, uint64_t n_out) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx || !ntt) return B200_ERR_INPUT_VALIDATION;
	if (log_len + log_batch > 60 || n_in != (1ull << (log_len + log_batch))) return fail(ctx, B200_ERR_INPUT_VALIDATION, "invalid data_in length: %llu", (unsigned long long)n_in);
	if (n_ch < log_batch) return fail(ctx, B200_ERR_INPUT_VALIDATION, "invalid challenges length: %u", n_ch);
	if (n_ch > log_batch + log_len) return fail(ctx, B200_ERR_INPUT_VALIDATION, "challenges length too big: %u", n_ch);
	uint32_t eta = n_ch - log_batch;
	if (n_out != (1ull << (log_len - eta))) return fail(ctx, B200_ERR_INPUT_VALIDATION, "invalid data_out length: %llu", (unsigned long long)n_out);
	if (eta > 0 && log_len > ntt->d) return fail(ctx, B200_ERR_NTT_DOMAIN, "domain size is less than 2**%u", log_len);
	if (n_ch > FRI_MAX_LOG_CHUNK) return fail(ctx, B200_ERR_INPUT_VALIDATION, "fri_fold: more than %u challenges per call not supported", FRI_MAX_LOG_CHUNK);
	std::vector<uint4> hc(std::max(n_ch, 1u));
	for (uint32_t i = 0; i < n_ch; i++) hc[i] = to_u4(challenges + 2 * i);
	void *dc;
	int32_t rc = stage_args(ctx, hc.data(), 16 * hc.size(), &dc);
	if (rc) return rc;
	FriArgs A{(const uint4 *)in, (uint4 *)out, n_out, log_len, log_batch, n_ch, (const uint4 *)dc, ntt->d_s_evals, ntt->d, ntt->kt};
	if (eta == 0 && n_ch >= 1 && n_ch <= 4 && n_out >= 4096) {
		// pure tensor-lerp fold (first FRI fold of the interleaved codeword): K64 tables per challenge
		const uint32_t nk = std::min(n_ch, 3u), smem = nk * LUT_BYTES + 6144 + NLUT_BYTES;
		uint32_t g = grid_for(ctx, n_out, FRI_K64_THREADS, 1);
		switch (n_ch) {
		case 1: k_fri_lerp_k64<1><<<g, FRI_K64_THREADS, smem, ctx->stream>>>(A); break;
		case 2: k_fri_lerp_k64<2><<<g, FRI_K64_THREADS, smem, ctx->stream>>>(A); break;
		case 3: k_fri_lerp_k64<3><<<g, FRI_K64_THREADS, smem, ctx->stream>>>(A); break;
		default: k_fri_lerp_k64<4><<<g, FRI_K64_THREADS, smem, ctx->stream>>>(A); break;
		}
		B200_LAUNCH_CHECK(ctx);
		return B200_OK;
	}
	// arities 1..5: register-resident chunk + nibble-LUT lerps (one 8 KiB table per challenge)
	uint32_t grid = grid_for(ctx, n_out, 256, 2);
	switch (n_ch) {
	case 1: k_fri_fold_lut<1><<<grid, 256, FIELD_TABLE_BYTES + 1 * NLUT_BYTES, ctx->stream>>>(ctx->d_tables, A); break;
	case 2: k_fri_fold_lut<2><<<grid, 256, FIELD_TABLE_BYTES + 2 * NLUT_BYTES, ctx->stream>>>(ctx->d_tables, A); break;
	case 3: k_fri_fold_lut<3><<<grid, 256, FIELD_TABLE_BYTES + 3 * NLUT_BYTES, ctx->stream>>>(ctx->d_tables, A); break;
	case 4: k_fri_fold_lut<4><<<grid, 256, FIELD_TABLE_BYTES + 4 * NLUT_BYTES, ctx->stream>>>(ctx->d_tables, A); break;
	case 5: k_fri_fold_lut<5><<<grid, 256, FIELD_TABLE_BYTES + 5 * NLUT_BYTES, ctx->stream>>>(ctx->d_tables, A); break;
	default: k_fri_fold<<<grid_for(ctx, n_out, 128, 2), 128, FIELD_TABLE_BYTES, ctx->stream>>>(ctx->d_tables, A); break;
	}
	B200_LAUNCH_CHECK(ctx);
	return B200_OK;
}

}  // extern "C"

// ---- zerocheck univariate-skip round (core/src/protocols/sumcheck/prove/univariate.rs:235-500) ---------
// Lagrange basis over the B8 points 0..n-1 at x (x not a node): L_t(x) = w_t * prod_u (x - u) / (x - t)
static void lagrange_b8(uint32_t n, uint32_t x, uint8_t *out) {
	const hostf::Tab8 &tb = hostf::tab8();
	auto mul = [&](uint32_t a, uint32_t b) -> uint32_t { return tb.mul[(a << 8) | b]; };
	// node weights w_t = 1 / prod_{u != t} (t - u): computed once per n (the host-side cost of a call used to be
	// O(points * n^2) table products, ~5 ms at n = 128, paid again by every chunk of a streamed round)
	static std::mutex mu;
	static std::map<uint32_t, std::vector<uint8_t>> weights;
	const uint8_t *w;
	{
		std::lock_guard<std::mutex> g(mu);
		auto it = weights.find(n);
		if (it == weights.end()) {
			std::vector<uint8_t> wv(n);
			for (uint32_t t = 0; t < n; t++) {
				uint32_t den = 1;
				for (uint32_t u = 0; u < n; u++)
					if (u != t) den = mul(den, t ^ u);
				wv[t] = (uint8_t)hostf::invert((hostf::u128)den, 3);
			}
			it = weights.emplace(n, std::move(wv)).first;
		}
		w = it->second.data();
	}
	uint32_t full = 1;
	for (uint32_t u = 0; u < n; u++) full = mul(full, x ^ u);
	for (uint32_t t = 0; t < n; t++) out[t] = (uint8_t)mul(mul(full, w[t]), (uint32_t)hostf::invert((hostf::u128)(x ^ t), 3));
}
static uint32_t const_level(uint64_t lo, uint64_t hi) {
	if (hi) return 7;
	if (lo >> 32) return 6;
	if (lo >> 16) return 5;
	if (lo >> 8) return 4;
	return 3;
}

// Validates and launches the round on the context's stream; *d_out_p = the [n_comp][n_out] device table (context scratch),
// nullptr when there is nothing to compute.  The kernels XOR into the table: `zero_out` clears it first, `extend` runs the
// (linear) domain extension afterwards -- a caller summing several sub-cube ranges clears once and extends once.  The caller copies it back (and may issue other copies first: the argument
// blocks of this call are already queued on the copy engine).
// The two halves of the B8 fast path (b200_zerocheck_univariate_prepare / _finish, univariate.cuh k_uni_finish): mode 1 runs
// k_uni_b8<SKIP, true> (values of the non-linear part of every composition to `store`, no challenge needed), mode 2 weights
// the stored values by eq and adds the linear-monomial route and the domain extension.
struct UniSplit {
	int mode = 0;            // 0: the whole round
	uint8_t *store = nullptr;
	uint64_t store_bytes = 0;
	uint64_t batch0 = 0;     // mode 1: first batch (of uni::SUBS sub-cubes) of this row chunk
	uint64_t n_eq_full = 0;  // mode 1: sub-cubes of the whole instance (the linear-monomial route is decided as mode 2 will)
	bool done = false;       // out: the fast path applied (otherwise nothing was launched)
};
static int32_t uni_issue(b200_ctx *ctx, const b200_dev_ptr *mls, const uint32_t *levels, uint32_t m, uint32_t n_vars, uint32_t skip, b200_dev_ptr eq_ind, uint64_t n_eq, const b200_expr *const *comps, const uint32_t *degrees, uint32_t n_comp, uint32_t max_domain_size, uint4 **d_out_p, bool zero_out = true, bool extend = true, UniSplit *sp = nullptr) {
	*d_out_p = nullptr;
	const int mode = sp ? sp->mode : 0;
	if (skip > n_vars) return fail(ctx, B200_ERR_INPUT_VALIDATION, "TooManySkippedRounds: skip_rounds %u > n_vars %u", skip, n_vars);
	if (n_vars - skip > 40) return fail(ctx, B200_ERR_INPUT_VALIDATION, "n_vars - skip_rounds must be <= 40");
	if (n_eq != (1ull << (n_vars - skip))) return fail(ctx, B200_ERR_INPUT_VALIDATION, "IncorrectZerocheckChallengesLength: eq_ind must hold 2^(n_vars - skip_rounds) elements");
	if (max_domain_size > 256) return fail(ctx, B200_ERR_INPUT_VALIDATION, "DomainSizeTooLarge: max_domain_size %u exceeds the B8 domain field", max_domain_size);
	if (m > uni::MAX_MLS || n_comp > uni::MAX_COMP) return fail(ctx, B200_ERR_INPUT_VALIDATION, "at most %u multilinears and %u compositions per call", uni::MAX_MLS, uni::MAX_COMP);
	uint32_t max_deg = 0, lvl = 3;
	for (uint32_t c = 0; c < n_comp; c++) {
		if (!comps[c] || comps[c]->n_vars > m) return fail(ctx, B200_ERR_INPUT_VALIDATION, "composition %u does not match the multilinears", c);
		max_deg = std::max(max_deg, degrees[c]);
		for (const b200_expr_step &st : comps[c]->steps)
			if (st.op == 3) lvl = std::max(lvl, const_level(st.c_lo, st.c_hi));
	}
	if (skip > 8 || (uint64_t)max_deg << skip > max_domain_size || max_domain_size < (1u << skip)) return fail(ctx, B200_ERR_INPUT_VALIDATION, "LagrangeDomainTooSmall: max_domain_size %u < %u << %u", max_domain_size, max_deg, skip);
	for (uint32_t j = 0; j < m; j++) {
		if (!valid_level(levels[j])) return fail(ctx, B200_ERR_INPUT_VALIDATION, "multilinear %u: unsupported tower level %u", j, levels[j]);
		if (!mls[j]) return fail(ctx, B200_ERR_INPUT_VALIDATION, "multilinear %u is null", j);
		lvl = std::max(lvl, levels[j]);
	}
	if (lvl == 6) lvl = 7;
	const uint32_t K = 1u << skip, n_out = max_domain_size - K;
	if (n_out == 0 || n_comp == 0) return B200_OK;
	const uint32_t n_pts = max_deg > 1 ? (max_deg - 1) << skip : 0;

	// scratch: the output table, then (linear-monomial route) E of up to m columns and the parity matrices of their jobs
	const uint64_t out_bytes = ((uint64_t)n_comp * n_out * 16 + 255) & ~255ull;
	int32_t rc = ensure_scratch(ctx, out_bytes + (uint64_t)m * 2048 + (uint64_t)(m + 1) * 2048);
	if (rc) return rc;
	uint4 *d_out = reinterpret_cast<uint4 *>(ctx->d_scratch);
	if (zero_out) B200_CUDA(ctx, cudaMemsetAsync(d_out, 0, (uint64_t)n_comp * n_out * 16, ctx->stream));
	if (n_pts) {
		std::vector<uint8_t> lag((size_t)n_pts * K);
		for (uint32_t i = 0; i < n_pts; i++) lagrange_b8(K, K + i, lag.data() + (size_t)i * K);
		std::vector<DevExpr> hc(n_comp);
		std::vector<uint32_t> pts(n_comp), ext_off(n_comp, 0);
		// extension matrices, one per distinct degree below the maximum
		std::vector<uint8_t> ext;
		std::map<uint32_t, uint32_t> ext_of_deg;
		bool need_ext = false;
		upload_exprs({{comps, n_comp}});
		for (uint32_t c = 0; c < n_comp; c++) {
			hc[c] = dev_expr(comps[c]);
			B200_EXPR_READY(ctx, hc[c]);
			pts[c] = degrees[c] > 1 ? (degrees[c] - 1) << skip : 0;
			if (pts[c] == 0 || pts[c] >= n_out) continue;
			need_ext = true;
			auto it = ext_of_deg.find(degrees[c]);
			if (it == ext_of_deg.end()) {
				const uint32_t n_nodes = degrees[c] << skip, n_in = pts[c];
				uint32_t off = (uint32_t)ext.size();
				ext.resize(off + (size_t)(n_out - n_in) * n_in);
				std::vector<uint8_t> row(n_nodes);
				for (uint32_t i = n_in; i < n_out; i++) {
					lagrange_b8(n_nodes, K + i, row.data());
					memcpy(ext.data() + off + (size_t)(i - n_in) * n_in, row.data() + K, n_in);
				}
				it = ext_of_deg.emplace(degrees[c], off).first;
			}
			ext_off[c] = it->second;
		}
		ArgPack pk;
		size_t o_mls = pk.add(mls, sizeof(void *) * m), o_lv = pk.add(levels, 4 * m), o_c = pk.add(hc.data(), sizeof(DevExpr) * n_comp);
		size_t o_p = pk.add(pts.data(), 4 * n_comp), o_l = pk.add(lag.data(), lag.size()), o_eo = pk.add(ext_off.data(), 4 * n_comp);
		size_t o_e = pk.add(ext.data(), ext.size());
		uint8_t *base;
		rc = pk.commit(ctx, &base);
		if (rc) return rc;
		uni::Args A;
		A.mls = (const uint4 *const *)(base + o_mls);
		A.levels = (const uint32_t *)(base + o_lv);
		A.comps = (const DevExpr *)(base + o_c);
		A.comp_pts = (const uint32_t *)(base + o_p);
		A.lag = base + o_l;
		A.eq = (const uint4 *)eq_ind;
		A.out = d_out;
		A.n_sub = n_eq;
		A.m = m, A.n_comp = n_comp, A.skip = skip, A.n_pts = n_pts, A.n_out = n_out;
		A.off_nl = (FIELD_TABLE_BYTES + K * K + 15) & ~15u;
		A.off_red = A.off_nl + (K >= 4 ? 4 * K * K : 0);
		const uint32_t smem = A.off_red + uni::THREADS * 16;
		const uint32_t SL = uni::THREADS >> skip, gy = n_pts >> skip;
		const uint32_t per_sm = std::max(1u, std::min(8u, (227u * 1024u) / (smem + 1024u)));
		const uint64_t want = (n_eq + SL - 1) / SL;
		const uint32_t gx = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(want, std::max(1u, ctx->n_sms * per_sm / gy)));
		dim3 grid(gx, gy);
		// B8 fast path: B1/B8 columns, B8 constants, every composition a sum of monomials of degree <= 2
		bool fast = lvl == 3 && skip >= 2 && !ctx->tune_uni_generic;
		std::vector<uint2> mono;
		std::vector<uint32_t> ctab(uni::CTAB * (size_t)n_comp);
		// coefficient-1 linear monomials leave the kernel when the sub-cube is 128 rows and the tensor-core outer product
		// applies (univariate.cuh, k_uni_linear)
		const uint64_t n_eq_route = mode == 1 ? sp->n_eq_full : n_eq;
		const bool lin_out = fast && skip == 7 && ctx->tune_uni_linear && ctx->tune_round_evals_tc && n_eq_route >= 8192 && n_eq_route % tc::CHUNK == 0;
		std::vector<uint32_t> lin_off(n_comp + 1, 0), lin_cols;
		std::map<uint32_t, uint32_t> lin_slot;  // column -> row of E
		for (uint32_t c = 0; fast && c < n_comp; c++) {
			if (!comps[c]->poly_ok) {
				fast = false;
				break;
			}
			// monomials in three runs: x_a * x_b and x_a with coefficient 1 (byte offsets into the staged
			// extrapolations), then everything else
			std::vector<uint2> quad, lin, gen;
			for (auto &t : comps[c]->poly) {
				if (t.second >> 8) fast = false;
				const size_t deg = t.first.size();
				if (t.second == 1 && deg == 2) quad.push_back(make_uint2(t.first[0] * uni::SUBS * K, t.first[1] * uni::SUBS * K));
				else if (t.second == 1 && deg == 1 && lin_out && levels[t.first[0]] == 0 && pts[c]) {
					auto it = lin_slot.emplace(t.first[0], (uint32_t)lin_slot.size()).first;
					lin_cols.push_back(it->second);
				} else if (t.second == 1 && deg == 1) lin.push_back(make_uint2(t.first[0] * uni::SUBS * K, 0));
				else gen.push_back(make_uint2((deg > 0 ? t.first[0] : uni::MONO_NONE) | ((deg > 1 ? t.first[1] : uni::MONO_NONE) << 9) | ((uint32_t)t.second << 18), 0));
			}
			uint32_t *ct = &ctab[uni::CTAB * c];
			ct[0] = (uint32_t)mono.size(), ct[1] = (uint32_t)quad.size(), ct[2] = (uint32_t)lin.size(), ct[3] = (uint32_t)gen.size(), ct[4] = pts[c];
			mono.insert(mono.end(), quad.begin(), quad.end());
			mono.insert(mono.end(), lin.begin(), lin.end());
			mono.insert(mono.end(), gen.begin(), gen.end());
			lin_off[c + 1] = (uint32_t)lin_cols.size();
		}
		const uint64_t rec_bytes = (uint64_t)n_comp * n_pts * uni::SUBS, n_batches_all = (n_eq + uni::SUBS - 1) / uni::SUBS;
		if (mode && fast && ((mode == 1 ? sp->batch0 : 0) + n_batches_all) * rec_bytes > sp->store_bytes)
			return fail(ctx, B200_ERR_INPUT_VALIDATION, "store holds %llu bytes, the prepared values need %llu", (unsigned long long)sp->store_bytes,
						(unsigned long long)(((mode == 1 ? sp->batch0 : 0) + n_batches_all) * rec_bytes));
		if (fast) {
			// shared-memory layout for `ml` columns / `nc` compositions / `nm` monomials; returns the dynamic size
			auto layout = [&](uni::B8Args &B, uint32_t ml, uint32_t nc, size_t nm) -> uint32_t {
				B.off_nl = A.off_nl;
				B.off_q = B.off_nl + 512 * K;  // (K/4 nibbles) x 16 patterns x 32-word rows
				B.off_es = B.off_q + ((ml * uni::SUBS * K + 15) & ~15u);
				B.off_mono = B.off_es + uni::SUBS * 32 * 16;
				B.off_ctab = B.off_mono + 8 * (uint32_t)std::max<size_t>(nm, 1);
				B.off_cols = (B.off_ctab + 4 * uni::CTAB * nc + 15) & ~15u;
				B.off_bits = (B.off_cols + 12 * ml + 15) & ~15u;
				return B.off_bits + ml * K + 16;  // ml * (SUBS * K / 32) words
			};
			auto fits = [&](uint32_t smem8, uint32_t ml) { return smem8 <= 227u * 1024u && ml * (uni::SUBS * K / 32) <= 8 * uni::B8_THREADS; };
			// one launch over compositions [c0, c0 + nc) (rows c0.. of the output table) and the given column list.
			// Two steps, so that a multi-range call stages ALL its argument blocks before its first kernel: an argument copy
			// queued between two kernels would wait behind whatever else occupies the host-to-device copy engine by then
			// (the next witness chunk of b200_zerocheck_univariate_evals_streamed) and stall the remaining ranges.
			auto stage = [&](uni::B8Args &B, const uint4 *const *d_mls, const uint32_t *d_levels, uint32_t ml, uint32_t c0, uint32_t nc,
							 const std::vector<uint2> &mn, const std::vector<uint32_t> &ct) -> int32_t {
				ArgPack pk2;
				size_t o_m = pk2.add(mn.data(), 8 * mn.size()), o_t = pk2.add(ct.data(), 4 * ct.size());
				uint8_t *base2;
				int32_t r = pk2.commit(ctx, &base2);
				if (r) return r;
				B.mls = d_mls, B.levels = d_levels, B.lag = A.lag, B.eq = A.eq, B.out = A.out + (uint64_t)c0 * n_out, B.n_sub = n_eq;
				B.mono = (const uint2 *)(base2 + o_m);
				B.comp_tab = (const uint32_t *)(base2 + o_t);
				B.m = ml, B.n_comp = nc, B.n_mono = (uint32_t)mn.size(), B.n_out = n_out;
				B.store = mode == 1 ? sp->store + sp->batch0 * rec_bytes + (uint64_t)c0 * n_pts * uni::SUBS : nullptr;
				B.rec_bytes = rec_bytes, B.n_pts = n_pts;
				return B200_OK;
			};
			auto launch = [&](const uni::B8Args &B, uint32_t smem8) -> int32_t {
				int32_t r;
				const uint64_t n_batches = (n_eq + uni::SUBS - 1) / uni::SUBS;
				dim3 grid8((uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_batches, std::max(1u, (uint32_t)ctx->n_sms / gy))), gy);
#define B200_UNI_B8(S)                                                                                    \
	case S:                                                                                               \
		if (mode == 1) {                                                                                  \
			if ((r = set_smem(ctx, uni::k_uni_b8<S, true>, smem8))) return r;                             \
			uni::k_uni_b8<S, true><<<grid8, uni::B8_THREADS, smem8, ctx->stream>>>(ctx->d_tables, B); \
		} else {                                                                                          \
			if ((r = set_smem(ctx, uni::k_uni_b8<S>, smem8))) return r;                                   \
			uni::k_uni_b8<S><<<grid8, uni::B8_THREADS, smem8, ctx->stream>>>(ctx->d_tables, B);       \
		}                                                                                                 \
		break;
				switch (skip) {
					B200_UNI_B8(2)
					B200_UNI_B8(3)
					B200_UNI_B8(4)
					B200_UNI_B8(5)
					B200_UNI_B8(6)
					B200_UNI_B8(7)
				}
#undef B200_UNI_B8
				return B200_OK;
			};
			uni::B8Args B;
			const uint32_t smem8 = layout(B, m, n_comp, mono.size());
			// (with linear monomials taken out some columns may not be referenced any more: the range planner compacts them away)
			if (lin_slot.empty() && fits(smem8, m)) {
				if (mode != 2 && ((rc = stage(B, A.mls, A.levels, m, 0, n_comp, mono, ctab)) || (rc = launch(B, smem8)))) return rc;
			} else {
				// Too many columns for one CTA's shared memory (e.g. 153 columns at skip 7): constraints are local,
				// so split the compositions into contiguous ranges whose referenced columns fit and launch the same
				// kernel per range on the compacted column list (verified on the GPU in round 1: both variants of
				// test_many_columns_at_the_reference_skip passed).
				static_assert(uni::SPLIT_MONO_NONE == uni::MONO_NONE && uni::SPLIT_CTAB == uni::CTAB && sizeof(uni::MonoW) == sizeof(uint2), "uni_split.hpp mirrors univariate.cuh");
				const uint32_t cube = uni::SUBS * K;
				std::vector<uni::MonoW> mw(mono.size());
				if (!mono.empty()) memcpy(mw.data(), mono.data(), 8 * mono.size());
				std::vector<uni::SplitRange> ranges;
				const size_t nm_all = mono.size();
				fast = uni::plan_split(mw, ctab, n_comp, m, cube, [&](uint32_t ncols, uint32_t ncomp) {
					uni::B8Args probe;
					return fits(layout(probe, ncols, ncomp, nm_all), ncols);
				}, ranges);
				std::vector<std::pair<uni::B8Args, uint32_t>> staged;
				for (size_t ri = 0; fast && mode != 2 && ri < ranges.size(); ri++) {  // (mode 2 plans only: same eligibility as mode 1)
					const uni::SplitRange &R = ranges[ri];
					std::vector<b200_dev_ptr> h_mls;
					std::vector<uint32_t> h_lv;
					for (uint32_t g : R.cols) {
						h_mls.push_back(mls[g]);
						h_lv.push_back(levels[g]);
					}
					if (h_mls.empty()) {  // constants only: the kernel still wants one (unused) column
						if (m == 0) {
							fast = false;
							break;
						}
						h_mls.push_back(mls[0]), h_lv.push_back(levels[0]);
					}
					std::vector<uint2> mn(R.mono.size());
					if (!mn.empty()) memcpy(mn.data(), R.mono.data(), 8 * mn.size());
					ArgPack pk3;
					size_t o_p = pk3.add(h_mls.data(), sizeof(void *) * h_mls.size()), o_l = pk3.add(h_lv.data(), 4 * h_lv.size());
					uint8_t *base3;
					if ((rc = pk3.commit(ctx, &base3))) return rc;
					uni::B8Args Br;
					const uint32_t ml = (uint32_t)h_mls.size();
					const uint32_t sm = layout(Br, ml, R.c1 - R.c0, mn.size());
					if ((rc = stage(Br, (const uint4 *const *)(base3 + o_p), (const uint32_t *)(base3 + o_l), ml, R.c0, R.c1 - R.c0, mn, R.ctab))) return rc;
					staged.emplace_back(Br, sm);
				}
				for (size_t ri = 0; fast && ri < staged.size(); ri++) {
					if ((rc = launch(staged[ri].first, staged[ri].second))) return rc;
					if (ri + 1 < staged.size()) B200_LAUNCH_CHECK(ctx);
				}
			}
			if (fast && mode == 2) {
				// the stored values weighted by eq: composition ranges sized to the two-deep ring and the per-thread accumulators
				const uint32_t buf_max = ((227u * 1024u - FIELD_TABLE_BYTES - uni::FIN_ES_BYTES - 4 * uni::MAX_COMP) / 2) & ~15u;
				const uint32_t range = std::max(1u, std::min({n_comp, (uni::FIN_ACC * uni::B8_THREADS) / n_pts, buf_max / (n_pts * uni::SUBS)}));
				if (range * n_pts > uni::FIN_ACC * uni::B8_THREADS || range * n_pts * uni::SUBS > buf_max) fast = false;
				else {
					uni::FinArgs FA;
					FA.store = sp->store, FA.comp_pts = A.comp_pts, FA.eq = A.eq, FA.out = d_out, FA.n_sub = n_eq, FA.rec_bytes = rec_bytes;
					FA.n_comp = n_comp, FA.n_pts = n_pts, FA.n_out = n_out, FA.range = range;
					FA.buf_bytes = (range * n_pts * uni::SUBS + 15) & ~15u;
					const uint32_t smem_f = FIELD_TABLE_BYTES + uni::FIN_ES_BYTES + 4 * uni::MAX_COMP + 2 * FA.buf_bytes;
					const uint32_t gy_f = (n_comp + range - 1) / range;
					if ((rc = set_smem(ctx, uni::k_uni_finish, smem_f))) return rc;
					dim3 grid_f((uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_batches_all, std::max(1u, (uint32_t)ctx->n_sms / gy_f))), gy_f);
					uni::k_uni_finish<<<grid_f, uni::B8_THREADS, smem_f, ctx->stream>>>(ctx->d_tables, FA);
				}
			}
			if (fast && !lin_slot.empty() && mode != 1) {
				B200_LAUNCH_CHECK(ctx);
				// E_j = evaluate_partial_high of every linearly used column by eq: one job per column, ONE tensor-core launch
				std::vector<tc::TcJob> jobs(lin_slot.size());
				for (auto &kv : lin_slot) jobs[kv.second] = tc::TcJob{(const uint4 *)eq_ind, nullptr, (const uint4 *)mls[kv.first], nullptr};
				const uint32_t n_slots = (uint32_t)lin_slot.size();
				uint4 *d_E = reinterpret_cast<uint4 *>(ctx->d_scratch + out_bytes);
				uint32_t *gmat = nullptr;
				if ((rc = launch_tc_pairs(ctx, jobs, n_eq, {}, out_bytes + (uint64_t)m * 2048, &gmat))) return rc;
				tc::k_tc_outer_combine<<<n_slots, 128, 0, ctx->stream>>>(gmat, 1, d_E);
				B200_LAUNCH_CHECK(ctx);
				ArgPack pk4;
				size_t o_lo = pk4.add(lin_off.data(), 4 * lin_off.size()), o_lc = pk4.add(lin_cols.data(), 4 * lin_cols.size());
				uint8_t *base4;
				if ((rc = pk4.commit(ctx, &base4))) return rc;
				uni::LinArgs LA{(const uint32_t *)(base4 + o_lo), (const uint32_t *)(base4 + o_lc), A.comp_pts, A.lag, d_E, d_out, n_out};
				if ((rc = set_smem(ctx, uni::k_uni_linear, FIELD_TABLE_BYTES))) return rc;
				uni::k_uni_linear<<<n_comp, 128, FIELD_TABLE_BYTES, ctx->stream>>>(ctx->d_tables, LA);
			}
		}
		if (mode) {
			sp->done = fast;
			if (!fast || mode == 1) {  // not a fast-path shape: nothing was launched, the caller runs the whole round instead
				B200_LAUNCH_CHECK(ctx);
				return B200_OK;
			}
		}
		if (!fast) switch (lvl) {
		case 3:
			if ((rc = set_smem(ctx, uni::k_uni_evals<3>, smem))) return rc;
			uni::k_uni_evals<3><<<grid, uni::THREADS, smem, ctx->stream>>>(ctx->d_tables, A);
			break;
		case 4:
			if ((rc = set_smem(ctx, uni::k_uni_evals<4>, smem))) return rc;
			uni::k_uni_evals<4><<<grid, uni::THREADS, smem, ctx->stream>>>(ctx->d_tables, A);
			break;
		case 5:
			if ((rc = set_smem(ctx, uni::k_uni_evals<5>, smem))) return rc;
			uni::k_uni_evals<5><<<grid, uni::THREADS, smem, ctx->stream>>>(ctx->d_tables, A);
			break;
		default:
			if ((rc = set_smem(ctx, uni::k_uni_evals<7>, smem))) return rc;
			uni::k_uni_evals<7><<<grid, uni::THREADS, smem, ctx->stream>>>(ctx->d_tables, A);
			break;
		}
		B200_LAUNCH_CHECK(ctx);
		if (need_ext && extend) {
			uni::ExtArgs X{A.comp_pts, (const uint32_t *)(base + o_eo), base + o_e, d_out, n_out};
			if ((rc = set_smem(ctx, uni::k_uni_extend, FIELD_TABLE_BYTES))) return rc;
			uni::k_uni_extend<<<n_comp, 256, FIELD_TABLE_BYTES, ctx->stream>>>(ctx->d_tables, X);
			B200_LAUNCH_CHECK(ctx);
		}
	}
	*d_out_p = d_out;
	return B200_OK;
}

int32_t b200_zerocheck_univariate_evals(b200_ctx *ctx, const b200_dev_ptr *mls, const uint32_t *levels, uint32_t m, uint32_t n_vars, uint32_t skip, b200_dev_ptr eq_ind, uint64_t n_eq, const b200_expr *const *comps, const uint32_t *degrees, uint32_t n_comp, uint32_t max_domain_size, uint64_t *host_out) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx) return B200_ERR_INPUT_VALIDATION;
	uint4 *d_out;
	int32_t rc = uni_issue(ctx, mls, levels, m, n_vars, skip, eq_ind, n_eq, comps, degrees, n_comp, max_domain_size, &d_out);
	if (rc || !d_out) return rc;
	if (!host_out) return B200_ERR_INPUT_VALIDATION;
	B200_CUDA(ctx, cudaMemcpyAsync(host_out, d_out, (uint64_t)n_comp * (max_domain_size - (1u << skip)) * 16, cudaMemcpyDeviceToHost, ctx->stream));
	B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return B200_OK;
}

// The same round with the witness still in HOST memory: column j (pinned, packed as on the device) is uploaded into mls[j]
// in 2^log_chunks row chunks on the side stream while the previous chunk is evaluated -- the round values are XOR-sums
// over sub-cubes and the domain extension is linear, so the chunks' tables add up.  The kernels of chunk c are issued
// BEFORE the copies of chunk c + 1: their (small) argument blocks travel through the same host-to-device copy engine
// and would otherwise queue behind 1/2^log_chunks of the witness.  Afterwards the columns are resident in mls[].
// (sp != nullptr: the challenge-independent half only -- b200_zerocheck_univariate_prepare -- with the same chunked upload)
static int32_t uni_streamed(b200_ctx *ctx, const void *const *host_cols, const b200_dev_ptr *mls, const uint32_t *levels, uint32_t m, uint32_t n_vars, uint32_t skip, b200_dev_ptr eq_ind, uint64_t n_eq, const b200_expr *const *comps, const uint32_t *degrees, uint32_t n_comp, uint32_t max_domain_size, uint32_t log_chunks, uint64_t *host_out, UniSplit *sp) {
	if (!ctx || !host_cols || !mls || !levels) return B200_ERR_INPUT_VALIDATION;
	if (skip > n_vars) return fail(ctx, B200_ERR_INPUT_VALIDATION, "TooManySkippedRounds: skip_rounds %u > n_vars %u", skip, n_vars);
	if (n_eq != (1ull << (n_vars - skip))) return fail(ctx, B200_ERR_INPUT_VALIDATION, "IncorrectZerocheckChallengesLength: eq_ind must hold 2^(n_vars - skip_rounds) elements");
	if (m == 0) return fail(ctx, B200_ERR_INPUT_VALIDATION, "NumberOfVariablesMismatch: no multilinears");
	log_chunks = std::min(log_chunks, n_vars - skip);
	for (uint32_t j = 0; j < m; j++) {
		if (!valid_level(levels[j])) return fail(ctx, B200_ERR_INPUT_VALIDATION, "multilinear %u: unsupported tower level %u", j, levels[j]);
		if (!host_cols[j] || !mls[j]) return fail(ctx, B200_ERR_INPUT_VALIDATION, "multilinear %u is null", j);
		while (log_chunks && (((1ull << (n_vars - log_chunks)) << levels[j]) & 127)) log_chunks--;  // a chunk of every column is whole B128 words
	}
	while (sp && log_chunks && ((n_eq >> log_chunks) % uni::SUBS)) log_chunks--;  // a chunk is whole batches of the value store
	int32_t rc = ensure_side_streams(ctx);
	if (rc) return rc;
	const uint32_t n_chunks = 1u << log_chunks;
	const uint64_t sub_per_chunk = n_eq >> log_chunks;
	const uint64_t n_vals = max_domain_size > (1u << skip) ? (uint64_t)n_comp * (max_domain_size - (1u << skip)) : 0;
	std::vector<uint64_t> bytes(m);
	for (uint32_t j = 0; j < m; j++) bytes[j] = std::max<uint64_t>(((1ull << (n_vars - log_chunks)) << levels[j]) / 8, 1);
	// Columns laid out at one stride on both sides (a witness arena, the usual case): a chunk of all columns is ONE
	// pitched copy.  One cudaMemcpyAsync per column and chunk (153 x 8 copies of 2 MiB for keccak 2^18) costs the copy
	// engine a few microseconds each, ~6 ms over the upload.
	bool pitched = m > 1;
	const ptrdiff_t hs = m > 1 ? (const uint8_t *)host_cols[1] - (const uint8_t *)host_cols[0] : 0, ds = m > 1 ? (const uint8_t *)mls[1] - (const uint8_t *)mls[0] : 0;
	for (uint32_t j = 0; j < m && pitched; j++)
		pitched = bytes[j] == bytes[0] && (const uint8_t *)host_cols[j] - (const uint8_t *)host_cols[0] == (ptrdiff_t)j * hs && (const uint8_t *)mls[j] - (const uint8_t *)mls[0] == (ptrdiff_t)j * ds;
	pitched = pitched && hs >= (ptrdiff_t)(bytes[0] << log_chunks) && ds >= (ptrdiff_t)(bytes[0] << log_chunks);
	auto upload = [&](uint32_t c) -> int32_t {
		if (pitched) {
			B200_CUDA(ctx, cudaMemcpy2DAsync((uint8_t *)mls[0] + c * bytes[0], (size_t)ds, (const uint8_t *)host_cols[0] + c * bytes[0], (size_t)hs, bytes[0], m, cudaMemcpyHostToDevice, ctx->s_h2d));
		} else {
			for (uint32_t j = 0; j < m; j++)
				B200_CUDA(ctx, cudaMemcpyAsync((uint8_t *)mls[j] + c * bytes[j], (const uint8_t *)host_cols[j] + c * bytes[j], bytes[j], cudaMemcpyHostToDevice, ctx->s_h2d));
		}
		B200_CUDA(ctx, cudaEventRecord(ctx->ev_in[c & 3], ctx->s_h2d));
		return B200_OK;
	};
	std::vector<b200_dev_ptr> ptrs(m);
	upload_exprs({{comps, n_comp}});
	for (uint32_t c = 0; c < n_comp; c++)  // the lazy device copies of the expressions go first, while the copy engine is idle
		if (comps[c]) {
			DevExpr de = dev_expr(comps[c]);
			B200_EXPR_READY(ctx, de);
		}
	struct SideFlag {  // argument blocks bypass the copy engine while the witness is in flight (stage_args)
		b200_ctx *c;
		explicit SideFlag(b200_ctx *x) : c(x) { c->side_uploads = true; }
		~SideFlag() { c->side_uploads = false; }
	} side_flag(ctx);
	if ((rc = upload(0))) return rc;
	uint4 *d_out = nullptr;
	// no host synchronisation inside the loop: the chunks XOR into one device table, so the host-side planning of chunk
	// c + 1 overlaps with the kernels of chunk c, and every copy is queued as early as it can be
	for (uint32_t c = 0; c < n_chunks; c++) {
		B200_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[c & 3], 0));
		for (uint32_t j = 0; j < m; j++) ptrs[j] = (b200_dev_ptr)((uint8_t *)mls[j] + c * bytes[j]);
		if (sp) {
			sp->batch0 = c * sub_per_chunk / uni::SUBS;
			rc = (c == 0 || sp->done) ? uni_issue(ctx, ptrs.data(), levels, m, n_vars - log_chunks, skip, nullptr, sub_per_chunk, comps, degrees, n_comp, max_domain_size, &d_out, false, false, sp) : B200_OK;
		} else
			rc = uni_issue(ctx, ptrs.data(), levels, m, n_vars - log_chunks, skip, (b200_dev_ptr)((uint4 *)eq_ind + c * sub_per_chunk), sub_per_chunk, comps, degrees, n_comp, max_domain_size, &d_out,
						   c == 0, c + 1 == n_chunks);
		if (rc) {
			cudaStreamSynchronize(ctx->s_h2d);
			cudaStreamSynchronize(ctx->stream);
			return rc;
		}
		if (c + 1 < n_chunks && (rc = upload(c + 1))) return rc;
	}
	if (d_out) {
		if (!host_out) return B200_ERR_INPUT_VALIDATION;
		B200_CUDA(ctx, cudaMemcpyAsync(host_out, d_out, 16 * n_vals, cudaMemcpyDeviceToHost, ctx->stream));
		B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	}
	B200_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[(n_chunks - 1) & 3], 0));
	return B200_OK;
}

int32_t b200_zerocheck_univariate_evals_streamed(b200_ctx *ctx, const void *const *host_cols, const b200_dev_ptr *mls, const uint32_t *levels, uint32_t m, uint32_t n_vars, uint32_t skip, b200_dev_ptr eq_ind, uint64_t n_eq, const b200_expr *const *comps, const uint32_t *degrees, uint32_t n_comp, uint32_t max_domain_size, uint32_t log_chunks, uint64_t *host_out) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	return uni_streamed(ctx, host_cols, mls, levels, m, n_vars, skip, eq_ind, n_eq, comps, degrees, n_comp, max_domain_size, log_chunks, host_out, nullptr);
}

// ---- the round in two halves: everything that needs no challenge first (while the witness is uploaded / committed) -------
uint64_t b200_zerocheck_univariate_store_elems(uint32_t n_vars, uint32_t skip, const uint32_t *degrees, uint32_t n_comp) {
	uint32_t max_deg = 0;
	for (uint32_t c = 0; degrees && c < n_comp; c++) max_deg = std::max(max_deg, degrees[c]);
	if (skip > n_vars || n_vars - skip > 40 || max_deg < 2 || skip > 8) return 0;
	const uint64_t n_batches = ((1ull << (n_vars - skip)) + uni::SUBS - 1) / uni::SUBS;
	return (n_batches * n_comp * ((uint64_t)(max_deg - 1) << skip) * uni::SUBS + 15) / 16;
}

int32_t b200_zerocheck_univariate_prepare(b200_ctx *ctx, const void *const *host_cols, const b200_dev_ptr *mls, const uint32_t *levels, uint32_t m, uint32_t n_vars, uint32_t skip, const b200_expr *const *comps, const uint32_t *degrees, uint32_t n_comp, uint32_t max_domain_size, uint32_t log_chunks, b200_dev_ptr store, uint64_t store_elems, uint32_t *prepared) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx || !mls || !levels || !prepared || (n_comp && (!comps || !degrees))) return B200_ERR_INPUT_VALIDATION;
	*prepared = 0;
	if (skip > n_vars) return fail(ctx, B200_ERR_INPUT_VALIDATION, "TooManySkippedRounds: skip_rounds %u > n_vars %u", skip, n_vars);
	if (m == 0) return fail(ctx, B200_ERR_INPUT_VALIDATION, "NumberOfVariablesMismatch: no multilinears");
	if (n_vars - skip > 40) return fail(ctx, B200_ERR_INPUT_VALIDATION, "n_vars - skip_rounds must be <= 40");
	const uint64_t n_eq = 1ull << (n_vars - skip);
	UniSplit sp;
	sp.mode = 1, sp.store = (uint8_t *)store, sp.store_bytes = store ? store_elems * 16 : 0, sp.n_eq_full = n_eq;
	int32_t rc;
	if (host_cols) rc = uni_streamed(ctx, host_cols, mls, levels, m, n_vars, skip, nullptr, n_eq, comps, degrees, n_comp, max_domain_size, log_chunks, nullptr, &sp);
	else {
		uint4 *d_out;
		rc = uni_issue(ctx, mls, levels, m, n_vars, skip, nullptr, n_eq, comps, degrees, n_comp, max_domain_size, &d_out, false, false, &sp);
	}
	if (rc) return rc;
	*prepared = sp.done ? 1 : 0;
	return B200_OK;
}

int32_t b200_zerocheck_univariate_finish(b200_ctx *ctx, const b200_dev_ptr *mls, const uint32_t *levels, uint32_t m, uint32_t n_vars, uint32_t skip, b200_dev_ptr eq_ind, uint64_t n_eq, const b200_expr *const *comps, const uint32_t *degrees, uint32_t n_comp, uint32_t max_domain_size, b200_dev_ptr store, uint64_t store_elems, uint64_t *host_out) {
	B200_LOCK(ctx);
	B200_FLUSH(ctx);
	if (!ctx || !mls || !levels || !store || !eq_ind) return B200_ERR_INPUT_VALIDATION;
	if (n_comp == 0 || max_domain_size == (1u << skip)) return B200_OK;
	UniSplit sp;
	sp.mode = 2, sp.store = (uint8_t *)store, sp.store_bytes = store_elems * 16;
	uint4 *d_out;
	int32_t rc = uni_issue(ctx, mls, levels, m, n_vars, skip, eq_ind, n_eq, comps, degrees, n_comp, max_domain_size, &d_out, true, true, &sp);
	if (rc) return rc;
	if (!sp.done) return fail(ctx, B200_ERR_INPUT_VALIDATION, "not a shape b200_zerocheck_univariate_prepare prepares (B1/B8 columns, B8 constants, monomials of degree <= 2, skip_rounds >= 2)");
	if (!d_out) return B200_OK;
	if (!host_out) return B200_ERR_INPUT_VALIDATION;
	B200_CUDA(ctx, cudaMemcpyAsync(host_out, d_out, (uint64_t)n_comp * (max_domain_size - (1u << skip)) * 16, cudaMemcpyDeviceToHost, ctx->stream));
	B200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return B200_OK;
}
