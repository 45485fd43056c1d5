// C++ conformance test of the host-side ComputeLayer mirror (binius_b200/host/compute_layer.hpp):
// replays scenarios of crates/compute_test_utils/src/layer.rs and the v3 bivariate round-evals
// program (core/src/protocols/sumcheck/v3/bivariate_product.rs:303-408) and compares with the CPU
// oracle (oracle/liboracle.so -- test infrastructure, linked by tests only).
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../binius_b200/host/compute_layer.hpp"
#include "../../binius_b200/host/computation_backend.hpp"

using namespace binius_b200;
typedef unsigned __int128 u128;
extern "C" {
int orc_extrapolate_line(void *e0, const void *e1, uint64_t n, const void *z);
int orc_tensor_expand(void *data, uint64_t len, uint32_t log_n, const void *coords, uint32_t k);
int orc_bivariate_round_evals(const void *const *mls, uint32_t m, uint32_t n_vars, const uint32_t *ia, const uint32_t *ib, uint32_t n_comp, const void *coeff, void *out2);
int orc_ntt_s_evals(uint32_t kt, uint32_t d, void *s);
int orc_ntt_transform(int inverse, const void *s, uint32_t kt, uint32_t d, void *data, uint32_t kd, uint64_t n, uint32_t lx, uint32_t ly, uint32_t lz, uint64_t coset, uint32_t cb, uint32_t skip);
void orc_mul(const void *a, const void *b, uint32_t k, void *out);
typedef struct { uint32_t op, l; uint64_t r; uint64_t c_lo, c_hi; } orc_expr_step;
int orc_sumcheck_round_evals(uint32_t order, const void *const *mls, const uint64_t *lens, const void *suffix, uint32_t m, uint32_t n_vars, const void *eq_ind,
							 const orc_expr_step *const *comps, const uint32_t *comp_steps, const orc_expr_step *const *leads, const uint32_t *lead_steps,
							 uint32_t n_comp, const uint32_t *codes, const void *points, uint32_t n_points, void *out);
uint64_t orc_fold_left_lerp_inplace(void *evals, uint64_t prefix, const void *suffix, uint32_t log_n, const void *z);
uint64_t orc_fold_right_lerp(const void *evals, uint64_t evals_size, const void *suffix, const void *z, void *out);
int orc_zerocheck_univariate_evals(const void *const *mls, const uint32_t *levels, uint32_t m, uint32_t n_vars, uint32_t skip, const void *eq_ind,
								   const orc_expr_step *const *comps, const uint32_t *comp_steps, uint32_t n_comp, uint32_t max_domain_size, void *out);
void orc_extrapolate_round_evals(uint32_t skip, uint32_t degree, uint32_t max_domain_size, void *vals);
}

static uint64_t sm_state;
static uint64_t splitmix() {
	uint64_t z = (sm_state += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
static std::vector<F128> rnd(uint64_t seed, size_t n) {
	sm_state = seed;
	std::vector<F128> v(n);
	for (auto &x : v) { x.lo = splitmix(); x.hi = splitmix(); }
	return v;
}
#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
	B200LayerHolder holder(1 << 10, 1 << 18);
	ComputeData data = holder.to_data();
	B200Layer &hal = data.hal;

	// extrapolate_line
	{
		size_t n = 1 << 12;
		auto e0 = rnd(1, n), e1 = rnd(2, n);
		F128 z = rnd(3, 1)[0];
		DevSlice d0 = data.dev_alloc.alloc(n), d1 = data.dev_alloc.alloc(n);
		hal.copy_h2d(e0.data(), n, d0);
		hal.copy_h2d(e1.data(), n, d1);
		hal.execute([&](B200Exec &ex) { ex.extrapolate_line(d0, d1, z); return std::vector<OpValue>{}; });
		std::vector<F128> got(n);
		hal.copy_d2h(d0, got.data(), n);
		orc_extrapolate_line(e0.data(), e1.data(), n, &z);
		CHECK(memcmp(got.data(), e0.data(), 16 * n) == 0);
		bool threw = false;
		try { hal.execute([&](B200Exec &ex) { ex.extrapolate_line(d0, d1.slice(0, n - 1), z); return std::vector<OpValue>{}; }); } catch (const InputValidation &) { threw = true; }
		CHECK(threw);
	}
	// tensor_expand
	{
		uint32_t log_n = 2, k = 9;
		std::vector<F128> h(1u << (log_n + k));
		auto v = rnd(4, 1u << log_n);
		for (size_t i = 0; i < v.size(); i++) h[i] = v[i];
		auto coords = rnd(5, k);
		DevSlice d = data.dev_alloc.alloc(h.size());
		hal.copy_h2d(h.data(), h.size(), d);
		hal.execute([&](B200Exec &ex) { ex.tensor_expand(log_n, coords, d); return std::vector<OpValue>{}; });
		std::vector<F128> got(h.size());
		hal.copy_d2h(d, got.data(), h.size());
		orc_tensor_expand(h.data(), h.size(), log_n, coords.data(), k);
		CHECK(memcmp(got.data(), h.data(), 16 * h.size()) == 0);
	}
	// v3 bivariate round evals as the traced accumulate_kernels program
	{
		uint32_t n_vars = 9, m = 4;
		size_t N = 1u << n_vars, half = N / 2;
		std::vector<std::vector<F128>> mls;
		std::vector<DevSlice> dml;
		for (uint32_t t = 0; t < m; t++) {
			mls.push_back(rnd(10 + t, N));
			dml.push_back(data.dev_alloc.alloc(N));
			hal.copy_h2d(mls[t].data(), N, dml[t]);
		}
		uint32_t ia[3] = {0, 2, 1}, ib[3] = {1, 3, 1};
		F128 alpha = rnd(20, 1)[0];
		std::vector<F128> pows{F128{1, 0}};
		for (int c = 0; c < 3; c++) { F128 nx; orc_mul(&pows.back(), &alpha, 7, &nx); pows.push_back(nx); }
		ExprEval prod = hal.compile_expr({ExprStep::var(0), ExprStep::var(1), ExprStep::mul(0, 1)});
		std::vector<KernelMemMap> maps;
		for (uint32_t t = 0; t < m; t++) {
			auto hs = dml[t].split_half();
			maps.push_back(KernelMemMap::chunked(hs.first, 0));
			maps.push_back(KernelMemMap::chunked(hs.second, 0));
			maps.push_back(KernelMemMap::local(n_vars - 1));
		}
		auto res = hal.execute([&](B200Exec &ex) {
			return ex.accumulate_kernels(
				[&](B200KernelExec &k, uint32_t log_chunks, std::vector<KernelBuffer> &b) {
					uint32_t sz = n_vars - 1 - log_chunks;
					OpValue y1 = k.decl_value(F128{});
					for (int c = 0; c < 3; c++) k.sum_composition_evals(SlicesBatch{{b[3 * ia[c] + 1].to_ref(), b[3 * ib[c] + 1].to_ref()}, 1ull << sz}, prod, pows[c], y1);
					for (uint32_t t = 0; t < m; t++) k.add(sz, b[3 * t].to_ref(), b[3 * t + 1].to_ref(), b[3 * t + 2].data);
					OpValue yi = k.decl_value(F128{});
					for (int c = 0; c < 3; c++) k.sum_composition_evals(SlicesBatch{{b[3 * ia[c] + 2].to_ref(), b[3 * ib[c] + 2].to_ref()}, 1ull << sz}, prod, pows[c], yi);
					return std::vector<OpValue>{y1, yi};
				},
				maps);
		});
		const void *ptrs[4] = {mls[0].data(), mls[1].data(), mls[2].data(), mls[3].data()};
		F128 exp[2];
		orc_bivariate_round_evals(ptrs, m, n_vars, ia, ib, 3, &alpha, exp);
		CHECK(res.size() == 2 && res[0] == exp[0] && res[1] == exp[1]);
		(void)half;
	}
	// AdditiveNTT on host data
	{
		B200Ntt ntt(hal, 5, 16);
		std::vector<F128> s(16 * 15);
		orc_ntt_s_evals(5, 16, s.data());
		sm_state = 77;
		std::vector<uint32_t> a(1 << 14), ref;
		for (auto &x : a) x = (uint32_t)splitmix();
		ref = a;
		ntt.forward_transform(a.data(), 5, a.size(), NTTShape{3, 11, 0}, 1, 2, 1);
		orc_ntt_transform(0, s.data(), 5, 16, ref.data(), 5, ref.size(), 3, 11, 0, 1, 2, 1);
		CHECK(a == ref);
		bool threw = false;
		try { ntt.forward_transform(a.data(), 5, a.size(), NTTShape{3, 11, 0}, 4, 2, 0); } catch (const NttError &e) { threw = e.code == B200_ERR_NTT_COSET; }
		CHECK(threw);
	}
	// ComputationBackend mirror: a regular (unweighted) sumcheck over 3 folded multilinears, composition
	// x*y + w at the points 1 and infinity, all rounds, in both evaluation orders, one truncated input
	for (int ord = 0; ord < 2; ord++) {
		const EvaluationOrder order = ord ? EvaluationOrder::HighToLow : EvaluationOrder::LowToHigh;
		B200Backend be(hal);
		const uint32_t n_vars = 9, m = 3;
		const uint64_t plen[3] = {1u << n_vars, (1u << n_vars) - 5, 300};
		F128 sfx[3] = {F128{}, rnd(90, 1)[0], F128{7, 0}};
		std::vector<std::vector<F128>> h(m);
		std::vector<SumcheckMultilinear> mls;
		for (uint32_t t = 0; t < m; t++) {
			h[t] = rnd(100 + t + 10 * ord, plen[t]);
			DevSlice d = hal.dev_alloc(plen[t]);
			hal.copy_h2d(h[t].data(), plen[t], d);
			mls.push_back(SumcheckMultilinear::folded(d, sfx[t]));
		}
		ExprEval comp = hal.compile_expr({ExprStep::var(0), ExprStep::var(1), ExprStep::mul(0, 1), ExprStep::var(2), ExprStep::add(2, 3)});
		ExprEval lead = hal.compile_expr({ExprStep::var(0), ExprStep::var(1), ExprStep::mul(0, 1)});
		const orc_expr_step oc[5] = {{4, 0, 0, 0, 0}, {4, 1, 0, 0, 0}, {1, 0, 1, 0, 0}, {4, 2, 0, 0, 0}, {0, 2, 3, 0, 0}};
		const orc_expr_step ol[3] = {{4, 0, 0, 0, 0}, {4, 1, 0, 0, 0}, {1, 0, 1, 0, 0}};
		const orc_expr_step *pc[1] = {oc}, *pl[1] = {ol};
		const uint32_t nc[1] = {5}, nl[1] = {3}, codes[2] = {1, 2};
		F128 zero2[2] = {};
		for (uint32_t rnd_i = 0; rnd_i < n_vars; rnd_i++) {
			const uint32_t nv = n_vars - rnd_i;
			SumcheckEvaluator ev{&comp, &lead, 1, 3};
			auto got = be.sumcheck_compute_round_evals(order, nv, nullptr, mls, {ev}, nullptr, {});
			const void *ptrs[3] = {h[0].data(), h[1].data(), h[2].data()};
			uint64_t lens[3] = {h[0].size(), h[1].size(), h[2].size()};
			F128 exp[2];
			orc_sumcheck_round_evals((uint32_t)order, ptrs, lens, sfx, m, nv, nullptr, pc, nc, pl, nl, 1, codes, zero2, 2, exp);
			CHECK(got.size() == 1 && got[0].size() == 2 && got[0][0] == exp[0] && got[0][1] == exp[1]);
			F128 ch = rnd(500 + rnd_i, 1)[0];
			CHECK(be.sumcheck_fold_multilinears(order, nv, mls, ch, nullptr) == false);
			for (uint32_t t = 0; t < m; t++) {
				if (ord) {
					h[t].resize(orc_fold_left_lerp_inplace(h[t].data(), h[t].size(), &sfx[t], nv, &ch));
				} else {
					std::vector<F128> o((h[t].size() + 1) / 2 + 1);
					o.resize(orc_fold_right_lerp(h[t].data(), h[t].size(), &sfx[t], &ch, o.data()));
					h[t] = o;
				}
				CHECK(mls[t].evals.n == h[t].size());
				std::vector<F128> back(h[t].size());
				hal.copy_d2h(mls[t].evals, back.data(), back.size());
				CHECK(memcmp(back.data(), h[t].data(), 16 * back.size()) == 0);
			}
		}
	}
	// zerocheck univariate-skip round (core/src/protocols/sumcheck/prove/univariate.rs:235-500): 3 B1 columns of 2^11
	// rows, skip 5, a degree-2 and a degree-3 composition, domain 3 * 2^5 + 7
	{
		B200Backend be(hal);
		const uint32_t n_vars = 11, skip = 5, max_domain = 103, n_out = max_domain - 32;
		std::vector<std::vector<F128>> h(3);
		std::vector<SumcheckMultilinear> mls;
		for (uint32_t t = 0; t < 3; t++) {
			h[t] = rnd(700 + t, 1u << (n_vars - 7));
			DevSlice d = hal.dev_alloc(h[t].size());
			hal.copy_h2d(h[t].data(), h[t].size(), d);
			mls.push_back(SumcheckMultilinear::transparent(d, 0, n_vars, 0));
		}
		ExprEval c2 = hal.compile_expr({ExprStep::var(0), ExprStep::var(1), ExprStep::mul(0, 1), ExprStep::var(2), ExprStep::add(2, 3)});
		ExprEval c3 = hal.compile_expr({ExprStep::var(0), ExprStep::var(1), ExprStep::mul(0, 1), ExprStep::var(2), ExprStep::mul(2, 3)});
		auto ch = rnd(710, n_vars - skip);
		auto out = zerocheck_univariate_evals(be, mls, {&c2, &c3}, {2, 3}, ch, skip, max_domain);
		CHECK(out.round_evals.size() == 2 && out.round_evals[0].size() == n_out && out.remaining_rounds == n_vars - skip);
		std::vector<F128> eq(1u << (n_vars - skip));
		eq[0] = F128{1, 0};
		orc_tensor_expand(eq.data(), eq.size(), 0, ch.data(), n_vars - skip);
		std::vector<F128> back(eq.size());
		hal.copy_d2h(out.partial_eq_ind_evals, back.data(), back.size());
		CHECK(memcmp(back.data(), eq.data(), 16 * eq.size()) == 0);
		const orc_expr_step o2[5] = {{4, 0, 0, 0, 0}, {4, 1, 0, 0, 0}, {1, 0, 1, 0, 0}, {4, 2, 0, 0, 0}, {0, 2, 3, 0, 0}};
		const orc_expr_step o3[5] = {{4, 0, 0, 0, 0}, {4, 1, 0, 0, 0}, {1, 0, 1, 0, 0}, {4, 2, 0, 0, 0}, {1, 2, 3, 0, 0}};
		const orc_expr_step *pc[2] = {o2, o3};
		const uint32_t nc[2] = {5, 5}, lv[3] = {0, 0, 0}, deg[2] = {2, 3};
		const void *ptrs[3] = {h[0].data(), h[1].data(), h[2].data()};
		std::vector<F128> exp(2 * n_out);
		CHECK(orc_zerocheck_univariate_evals(ptrs, lv, 3, n_vars, skip, eq.data(), pc, nc, 2, max_domain, exp.data()) == 0);
		for (uint32_t c = 0; c < 2; c++) {
			orc_extrapolate_round_evals(skip, deg[c], max_domain, exp.data() + c * n_out);
			CHECK(memcmp(out.round_evals[c].data(), exp.data() + c * n_out, 16 * n_out) == 0);
		}
		bool threw = false;
		try { zerocheck_univariate_evals(be, mls, {&c2, &c3}, {2, 3}, ch, skip, 95); } catch (const InputValidation &) { threw = true; }
		CHECK(threw);  // LagrangeDomainTooSmall
	}
	bool oom = false;
	try { data.dev_alloc.alloc(1 << 20); } catch (const AllocError &) { oom = true; }
	CHECK(oom);
	printf("cpp conformance ok\n");
	return 0;
}
