// computation_backend.hpp -- C++ host-side mirror of the reference's old HAL,
// `binius_hal::ComputationBackend` (crates/hal/src/backend.rs:35-83), over the C ABI, for
// device-resident multilinears.  Same names and argument meaning as the trait:
//   tensor_product_full_query      backend.rs:42-45
//   sumcheck_compute_round_evals   backend.rs:48-62  (sumcheck_round_calculation.rs:126-349)
//   sumcheck_fold_multilinears     backend.rs:65-75  (sumcheck_folding.rs:37-262)
//   evaluate_partial_high          backend.rs:78-82
// plus fold_partial_eq_ind (core/src/protocols/sumcheck/prove/common.rs:13-73), which the provers
// call next to the backend.  `SumcheckMultilinear` (hal/src/sumcheck_multilinear.rs:8-40) keeps both
// variants; a Transparent multilinear is a packed sub-field multilinear on the device that is
// partially evaluated by the tensor-query expansion (fold_left / fold_right at its tower level) per
// round until its switchover round, then replaced by the Folded result.
#pragma once
#include <algorithm>
#include <map>
#include <memory>

#include "compute_layer.hpp"

namespace binius_b200 {

enum class EvaluationOrder : uint32_t { LowToHigh = B200_LOW_TO_HIGH, HighToLow = B200_HIGH_TO_LOW };

struct SumcheckMultilinear {
	enum Kind { Transparent, Folded } kind = Folded;
	// Folded { large_field_folded_evals, suffix_eval }: stored prefix, the rest equals suffix_eval
	DevSlice evals;
	F128 suffix_eval;
	// Transparent { multilinear, switchover_round, const_suffix }: 2^n_vars scalars of tower level `tower_level`
	uint32_t tower_level = 7, n_vars = 0, switchover_round = 0;
	uint64_t suffix_len = 0;

	static SumcheckMultilinear folded(DevSlice evals, F128 suffix = {}) {
		SumcheckMultilinear m;
		m.kind = Folded; m.evals = evals; m.suffix_eval = suffix;
		return m;
	}
	static SumcheckMultilinear transparent(DevSlice packed, uint32_t tower_level, uint32_t n_vars, uint32_t switchover_round) {
		SumcheckMultilinear m;
		m.kind = Transparent; m.evals = packed; m.tower_level = tower_level; m.n_vars = n_vars; m.switchover_round = switchover_round;
		return m;
	}
};

// What reaches the backend of a SumcheckEvaluator: the composition (and its leading term, used at the
// infinity point), the evaluation-point range, and whether sums are weighted by eq_ind_partial_eval().
struct SumcheckEvaluator {
	const ExprEval *composition = nullptr;
	const ExprEval *composition_at_infinity = nullptr;
	uint32_t first_point = 1, end_point = 2;  // eval_point_indices(): 1 = at 1, 2 = infinity, k >= 3 finite
};

class B200Backend {
	B200Layer &l_;
	DevSlice unit_{}, ones2_{};
	// Out-of-place results of the LowToHigh paths (fold_right_lerp, eq-ind halving) come from this pool and the
	// buffer they replace goes back to it, so two buffers per multilinear ping-pong instead of one allocation
	// (= one device synchronisation) per round.  Buffers the caller handed in are never recycled; the pool is
	// released with the backend.
	// persistent sumcheck tail (b200_sumcheck_tail_*), opt-in: while it runs the layer accepts no other call, so only
	// a caller whose round loop does nothing else on the layer (the prover's) should switch it on
	uint64_t tail_threshold_ = 0;
	b200_tail *tail_ = nullptr;
	uint32_t tail_vars_ = 0, tail_eq_pending_ = 0, tail_hi_ = 0;
	bool tail_first_ = false;
	uint8_t *tail_eq_ptr_ = nullptr;
	std::map<uint8_t *, uint64_t> mine_;
	std::vector<std::pair<uint8_t *, uint64_t>> free_, all_;
	DevSlice take(uint64_t n) {
		n = std::max<uint64_t>(n, 1);
		size_t best = free_.size();
		for (size_t i = 0; i < free_.size(); i++)
			if (free_[i].second >= n && (best == free_.size() || free_[i].second < free_[best].second)) best = i;
		uint8_t *p;
		uint64_t cap;
		if (best != free_.size()) {
			p = free_[best].first, cap = free_[best].second;
			free_.erase(free_.begin() + best);
		} else {
			p = l_.dev_alloc(n).ptr, cap = n;
			all_.push_back({p, cap});
		}
		mine_[p] = cap;
		return DevSlice{p, n};
	}
	void give(DevSlice s) {
		auto it = mine_.find(s.ptr);
		if (it == mine_.end()) return;
		free_.push_back({it->first, it->second});  // one in-order stream: a later writer runs after the last reader
		mine_.erase(it);
	}

	DevSlice one() {
		if (!unit_.n) {
			unit_ = l_.dev_alloc(1);
			l_.fill(unit_, F128{1, 0});
		}
		return unit_;
	}
	DevSlice partial_eval(EvaluationOrder order, const SumcheckMultilinear &m, const DevSlice *tensor_query) {
		DevSlice q = tensor_query ? *tensor_query : one();
		const uint64_t n_scalars = 1ull << m.n_vars;
		if (!q.n || n_scalars % q.n) throw InputValidation(1, "tensor query does not divide the multilinear");
		DevSlice out = l_.dev_alloc(n_scalars / q.n);
		auto fn = order == EvaluationOrder::HighToLow ? b200_fold_left : b200_fold_right;
		l_.check(fn(l_.ctx(), m.evals.ptr, m.evals.n, m.tower_level, q.ptr, q.n, out.ptr, out.n));
		return out;
	}

  public:
	explicit B200Backend(B200Layer &l, bool sumcheck_tail = false) : l_(l), tail_threshold_(sumcheck_tail ? (1u << 20) : 0) {}
	~B200Backend() {
		for (auto &b : all_) b200_dev_free(l_.ctx(), b.first);
	}
	B200Backend(const B200Backend &) = delete;

	DevSlice tensor_product_full_query(const std::vector<F128> &query) {
		DevSlice out = l_.dev_alloc(1ull << query.size());
		l_.check(b200_tensor_product_full_query(l_.ctx(), (const uint64_t *)query.data(), (uint32_t)query.size(), out.ptr, out.n));
		return out;
	}

	// RoundEvals per evaluator (values at its eval_point_indices()).  eq_ind_partial_evals = nullptr: regular evaluator.
	std::vector<std::vector<F128>> sumcheck_compute_round_evals(EvaluationOrder order, uint32_t n_vars, const DevSlice *tensor_query,
																const std::vector<SumcheckMultilinear> &multilinears,
																const std::vector<SumcheckEvaluator> &evaluators, const DevSlice *eq_ind_partial_evals,
																const std::vector<F128> &nontrivial_evaluation_points) {
		if (n_vars == 0) throw InputValidation(1, "Computing round evaluations requires at least a single variable.");
		uint32_t lo = ~0u, hi = 0;
		for (auto &e : evaluators) { lo = std::min(lo, e.first_point); hi = std::max(hi, e.end_point); }
		if (evaluators.empty() || lo >= hi) return std::vector<std::vector<F128>>(evaluators.size());
		if (nontrivial_evaluation_points.size() != (hi > 3 ? hi - 3 : 0)) throw InputValidation(1, "IncorrectNontrivialEvalPointsLength");
		if (eq_ind_partial_evals && eq_ind_partial_evals->n != (1ull << (n_vars - 1))) throw InputValidation(1, "eq_ind_partial_evals must have 2^(n_vars-1) elements");
		std::vector<uint32_t> codes;
		std::vector<F128> pts;
		for (uint32_t c = lo; c < hi; c++) { codes.push_back(c); pts.push_back(c < 3 ? F128{} : nontrivial_evaluation_points[c - 3]); }
		// ---- persistent tail: start it when the round is small enough, then only read its mailbox.  The tail's point
		//      list is 1..hi-1 in every round; a prover's first round starts at lo = 2 (first_round_skip = lo - 1).
		const uint32_t n_tail_codes = hi - 1;
		bool tail_ok = tail_ != nullptr;
		if (!tail_ok && tail_threshold_ && eq_ind_partial_evals && order == EvaluationOrder::HighToLow && n_vars <= 28) {
			const uint64_t n_vals = (uint64_t)evaluators.size() * n_tail_codes;
			const bool grid = n_vars - 1 > 5;  // more than 32 hypercube points: the co-resident grid kernel
			tail_ok = (n_vals << (n_vars - 1)) <= tail_threshold_ && (grid ? n_vals <= 1024 : (n_vals <= 4096 && lo == 1));
			for (auto &m : multilinears) tail_ok = tail_ok && m.kind == SumcheckMultilinear::Folded && m.evals.n == (1ull << n_vars);
			if (tail_ok) {
				std::vector<b200_dev_ptr> p;
				for (auto &m : multilinears) p.push_back(m.evals.ptr);
				std::vector<const b200_expr *> cs, ls;
				for (auto &e : evaluators) { cs.push_back(e.composition->raw()); ls.push_back(e.composition_at_infinity->raw()); }
				std::vector<uint32_t> tc;
				std::vector<F128> tp;
				for (uint32_t c = 1; c < hi; c++) { tc.push_back(c); tp.push_back(c < 3 ? F128{} : nontrivial_evaluation_points[c - 3]); }
				l_.check(b200_sumcheck_tail_start(l_.ctx(), p.data(), (uint32_t)p.size(), n_vars, eq_ind_partial_evals->ptr, cs.data(), ls.data(),
												  (uint32_t)cs.size(), tc.data(), (const uint64_t *)tp.data(), n_tail_codes, lo - 1, &tail_));
				tail_vars_ = n_vars, tail_hi_ = hi, tail_first_ = true, tail_eq_ptr_ = eq_ind_partial_evals->ptr, tail_eq_pending_ = 0;
			}
		}
		if (tail_ok) {
			if (tail_vars_ != n_vars || tail_hi_ != hi || (lo != 1 && !tail_first_)) throw InputValidation(1, "the running sumcheck tail was started for a different round");
			tail_first_ = false;
			std::vector<F128> vals(evaluators.size() * n_tail_codes);
			l_.check(b200_sumcheck_tail_round_evals(tail_, (uint64_t *)vals.data()));
			std::vector<std::vector<F128>> res;
			for (size_t e = 0; e < evaluators.size(); e++) {
				std::vector<F128> r;
				for (uint32_t k = evaluators[e].first_point; k < evaluators[e].end_point; k++) r.push_back(vals[e * n_tail_codes + (k - 1)]);
				res.push_back(r);
			}
			return res;
		}
		std::vector<DevSlice> temps;
		std::vector<b200_dev_ptr> ptrs;
		std::vector<uint64_t> lens;
		std::vector<F128> sfx;
		for (auto &m : multilinears) {
			DevSlice v = m.evals;
			F128 s = m.suffix_eval;
			if (m.kind == SumcheckMultilinear::Transparent) {
				v = partial_eval(order, m, tensor_query);
				if (v.n != (1ull << n_vars)) throw InputValidation(1, "transparent multilinear does not match n_vars and the tensor query");
				temps.push_back(v);
				if (!m.suffix_len) s = F128{};
			}
			ptrs.push_back(v.ptr);
			lens.push_back(std::min<uint64_t>(v.n, 1ull << n_vars));
			sfx.push_back(s);
		}
		std::vector<const b200_expr *> comps, leads;
		for (auto &e : evaluators) { comps.push_back(e.composition->raw()); leads.push_back(e.composition_at_infinity->raw()); }
		uint32_t first = 0;
		l_.check(b200_results_reset(l_.ctx()));
		l_.check(b200_sumcheck_round_evals(l_.ctx(), (uint32_t)order, ptrs.data(), lens.data(), (const uint64_t *)sfx.data(), (uint32_t)ptrs.size(), n_vars,
										   eq_ind_partial_evals ? eq_ind_partial_evals->ptr : nullptr, comps.data(), leads.data(), (uint32_t)comps.size(),
										   codes.data(), (const uint64_t *)pts.data(), (uint32_t)codes.size(), &first));
		const uint32_t total = (uint32_t)(comps.size() * codes.size());
		std::vector<uint32_t> slots(total);
		for (uint32_t i = 0; i < total; i++) slots[i] = first + i;
		std::vector<F128> vals(total);
		l_.check(b200_results_fetch(l_.ctx(), slots.data(), total, (uint64_t *)vals.data()));  // synchronises
		for (auto &t : temps) l_.dev_free(t);
		std::vector<std::vector<F128>> res;
		for (size_t e = 0; e < evaluators.size(); e++) {
			std::vector<F128> r;
			for (uint32_t k = evaluators[e].first_point; k < evaluators[e].end_point; k++) r.push_back(vals[e * codes.size() + (k - lo)]);
			res.push_back(r);
		}
		return res;
	}

	// returns any_transparent_left; `tensor_query` already includes `challenge` (prover_state.rs:161-181)
	bool sumcheck_fold_multilinears(EvaluationOrder order, uint32_t n_vars, std::vector<SumcheckMultilinear> &multilinears, F128 challenge,
									const DevSlice *tensor_query) {
		if (tail_) {
			if (tail_vars_ != n_vars) throw InputValidation(1, "fold does not match the running sumcheck tail");
			uint64_t z[2] = {challenge.lo, challenge.hi};
			l_.check(b200_sumcheck_tail_challenge(tail_, z));
			for (auto &m : multilinears) m.evals = m.evals.slice(0, 1ull << (n_vars - 1));
			tail_eq_pending_++;
			if (--tail_vars_ == 0) {
				b200_tail *t = tail_;
				tail_ = nullptr;
				l_.check(b200_sumcheck_tail_finish(t));
			}
			return false;
		}
		bool any_transparent_left = false;
		std::vector<SumcheckMultilinear *> folded;
		for (auto &m : multilinears) {
			if (m.kind == SumcheckMultilinear::Transparent) {
				if (m.switchover_round == 0) {
					if (!tensor_query) throw InputValidation(1, "tensor_query is required while a multilinear is transparent");
					m = SumcheckMultilinear::folded(partial_eval(order, m, tensor_query), m.suffix_len ? m.suffix_eval : F128{});
				} else {
					m.switchover_round--;
					any_transparent_left = true;
				}
			} else {
				folded.push_back(&m);
			}
		}
		if (folded.empty()) return any_transparent_left;
		std::vector<b200_dev_ptr> ptrs, outs;
		std::vector<uint64_t> prefix, new_lens(folded.size());
		std::vector<F128> sfx;
		std::vector<DevSlice> out_slices;
		for (auto *m : folded) {
			ptrs.push_back(m->evals.ptr);
			prefix.push_back(std::min<uint64_t>(m->evals.n, 1ull << n_vars));
			sfx.push_back(m->suffix_eval);
		}
		uint64_t z[2] = {challenge.lo, challenge.hi};
		if (order == EvaluationOrder::HighToLow) {
			l_.check(b200_fold_multilinears_high_to_low(l_.ctx(), ptrs.data(), (uint32_t)ptrs.size(), n_vars, prefix.data(), (const uint64_t *)sfx.data(), z, new_lens.data()));
			for (size_t t = 0; t < folded.size(); t++) folded[t]->evals = folded[t]->evals.slice(0, new_lens[t]);
		} else {
			for (auto p : prefix) {
				out_slices.push_back(take((p + 1) / 2));
				outs.push_back(out_slices.back().ptr);
			}
			l_.check(b200_fold_multilinears_low_to_high(l_.ctx(), ptrs.data(), outs.data(), (uint32_t)ptrs.size(), n_vars, prefix.data(), (const uint64_t *)sfx.data(), z, new_lens.data()));
			for (size_t t = 0; t < folded.size(); t++) {
				give(folded[t]->evals);
				folded[t]->evals = out_slices[t].slice(0, new_lens[t]);
			}
		}
		return any_transparent_left;
	}

	DevSlice evaluate_partial_high(DevSlice multilinear, DevSlice query_expansion) {
		if (!query_expansion.n || multilinear.n % query_expansion.n) throw InputValidation(1, "query expansion must divide the multilinear");
		DevSlice out = l_.dev_alloc(multilinear.n / query_expansion.n);
		l_.check(b200_fold_left(l_.ctx(), multilinear.ptr, multilinear.n, 7, query_expansion.ptr, query_expansion.n, out.ptr, out.n));
		return out;
	}

	DevSlice fold_partial_eq_ind(EvaluationOrder order, uint32_t n_vars, DevSlice eq_ind) {
		if (n_vars == 0) return eq_ind;
		if (tail_ && tail_eq_ptr_ == eq_ind.ptr && tail_eq_pending_) {  // halved in place by the tail kernel
			tail_eq_pending_--;
			return eq_ind.slice(0, 1ull << (n_vars - 1));
		}
		if (order == EvaluationOrder::LowToHigh) {
			if (!ones2_.n) {
				ones2_ = l_.dev_alloc(2);
				l_.fill(ones2_, F128{1, 0});
			}
			DevSlice out = take(1ull << (n_vars - 1));
			l_.check(b200_fold_right(l_.ctx(), eq_ind.ptr, eq_ind.n, 7, ones2_.ptr, 2, out.ptr, out.n));
			give(eq_ind);
			return out;
		}
		auto halves = eq_ind.split_half();
		l_.check(b200_kernel_add(l_.ctx(), n_vars - 1, halves.first.ptr, halves.second.ptr, halves.first.ptr));
		return halves.first;
	}

	B200Layer &layer() { return l_; }
};

// core/src/protocols/sumcheck/prove/univariate.rs:37-50
struct ZerocheckUnivariateEvalsOutput {
	std::vector<std::vector<F128>> round_evals;
	uint32_t skip_rounds = 0, remaining_rounds = 0, max_domain_size = 0;
	DevSlice partial_eq_ind_evals;
};

// zerocheck_univariate_evals (core/src/protocols/sumcheck/prove/univariate.rs:235-500), FDomain = BinaryField8b.
// `multilinears`: Transparent (packed sub-field) multilinears of equal n_vars; `composition_degrees[c]` =
// CompositionPoly::degree() of compositions[c].  Same checks and error order as the reference.
inline ZerocheckUnivariateEvalsOutput zerocheck_univariate_evals(B200Backend &backend, const std::vector<SumcheckMultilinear> &multilinears,
																   const std::vector<const ExprEval *> &compositions,
																   const std::vector<uint32_t> &composition_degrees,
																   const std::vector<F128> &zerocheck_challenges, uint32_t skip_rounds,
																   uint32_t max_domain_size) {
	if (multilinears.empty()) throw InputValidation(1, "NumberOfVariablesMismatch: no multilinears");
	const uint32_t n_vars = multilinears[0].n_vars;
	for (auto &m : multilinears)
		if (m.kind != SumcheckMultilinear::Transparent || m.n_vars != n_vars) throw InputValidation(1, "NumberOfVariablesMismatch");
	if (skip_rounds > n_vars) throw InputValidation(1, "TooManySkippedRounds");
	if (zerocheck_challenges.size() != n_vars - skip_rounds) throw InputValidation(1, "IncorrectZerocheckChallengesLength");
	if (compositions.size() != composition_degrees.size()) throw InputValidation(1, "one degree per composition");
	uint32_t max_deg = 0;
	for (uint32_t d : composition_degrees) max_deg = std::max(max_deg, d);
	if ((uint64_t)max_domain_size < ((uint64_t)max_deg << skip_rounds)) throw InputValidation(1, "LagrangeDomainTooSmall");
	if (max_domain_size > 256) throw InputValidation(1, "DomainSizeTooLarge");
	ZerocheckUnivariateEvalsOutput out;
	out.skip_rounds = skip_rounds, out.remaining_rounds = n_vars - skip_rounds, out.max_domain_size = max_domain_size;
	out.partial_eq_ind_evals = backend.tensor_product_full_query(zerocheck_challenges);
	std::vector<b200_dev_ptr> ptrs;
	std::vector<uint32_t> levels;
	for (auto &m : multilinears) { ptrs.push_back(m.evals.ptr); levels.push_back(m.tower_level); }
	std::vector<const b200_expr *> exprs;
	for (auto *c : compositions) exprs.push_back(c->raw());
	const uint32_t n_out = max_domain_size - (1u << skip_rounds);
	std::vector<F128> flat(std::max<size_t>(compositions.size() * n_out, 1));
	B200Layer &l = backend.layer();
	l.check(b200_zerocheck_univariate_evals(l.ctx(), ptrs.data(), levels.data(), (uint32_t)ptrs.size(), n_vars, skip_rounds,
											out.partial_eq_ind_evals.ptr, out.partial_eq_ind_evals.n, exprs.data(), composition_degrees.data(),
											(uint32_t)exprs.size(), max_domain_size, (uint64_t *)flat.data()));
	for (size_t c = 0; c < compositions.size(); c++) out.round_evals.emplace_back(flat.begin() + c * n_out, flat.begin() + (c + 1) * n_out);
	return out;
}

// The same round with the witness columns still in (pinned) HOST memory: b200_zerocheck_univariate_evals_streamed uploads
// host_columns[j] into multilinears[j].evals chunk by chunk behind the evaluation of the previous chunk; afterwards the
// columns are device-resident for the multilinear rounds.  Values are identical to copy_h2d + zerocheck_univariate_evals.
inline ZerocheckUnivariateEvalsOutput zerocheck_univariate_evals_streamed(B200Backend &backend, const std::vector<const void *> &host_columns,
																		   const std::vector<SumcheckMultilinear> &multilinears,
																		   const std::vector<const ExprEval *> &compositions,
																		   const std::vector<uint32_t> &composition_degrees,
																		   const std::vector<F128> &zerocheck_challenges, uint32_t skip_rounds,
																		   uint32_t max_domain_size, uint32_t log_chunks = 3) {
	if (multilinears.empty() || host_columns.size() != multilinears.size()) throw InputValidation(1, "NumberOfVariablesMismatch: one host column per multilinear");
	const uint32_t n_vars = multilinears[0].n_vars;
	for (auto &m : multilinears)
		if (m.kind != SumcheckMultilinear::Transparent || m.n_vars != n_vars) throw InputValidation(1, "NumberOfVariablesMismatch");
	if (skip_rounds > n_vars) throw InputValidation(1, "TooManySkippedRounds");
	if (zerocheck_challenges.size() != n_vars - skip_rounds) throw InputValidation(1, "IncorrectZerocheckChallengesLength");
	if (compositions.size() != composition_degrees.size()) throw InputValidation(1, "one degree per composition");
	uint32_t max_deg = 0;
	for (uint32_t d : composition_degrees) max_deg = std::max(max_deg, d);
	if ((uint64_t)max_domain_size < ((uint64_t)max_deg << skip_rounds)) throw InputValidation(1, "LagrangeDomainTooSmall");
	if (max_domain_size > 256) throw InputValidation(1, "DomainSizeTooLarge");
	ZerocheckUnivariateEvalsOutput out;
	out.skip_rounds = skip_rounds, out.remaining_rounds = n_vars - skip_rounds, out.max_domain_size = max_domain_size;
	out.partial_eq_ind_evals = backend.tensor_product_full_query(zerocheck_challenges);
	std::vector<b200_dev_ptr> ptrs;
	std::vector<uint32_t> levels;
	for (auto &m : multilinears) { ptrs.push_back(m.evals.ptr); levels.push_back(m.tower_level); }
	std::vector<const b200_expr *> exprs;
	for (auto *c : compositions) exprs.push_back(c->raw());
	const uint32_t n_out = max_domain_size - (1u << skip_rounds);
	std::vector<F128> flat(std::max<size_t>(compositions.size() * n_out, 1));
	B200Layer &l = backend.layer();
	l.check(b200_zerocheck_univariate_evals_streamed(l.ctx(), host_columns.data(), ptrs.data(), levels.data(), (uint32_t)ptrs.size(), n_vars, skip_rounds,
													 out.partial_eq_ind_evals.ptr, out.partial_eq_ind_evals.n, exprs.data(), composition_degrees.data(),
													 (uint32_t)exprs.size(), max_domain_size, log_chunks, (uint64_t *)flat.data()));
	for (size_t c = 0; c < compositions.size(); c++) out.round_evals.emplace_back(flat.begin() + c * n_out, flat.begin() + (c + 1) * n_out);
	return out;
}

// The round in two halves (b200_zerocheck_univariate_prepare / _finish): `zerocheck_univariate_prepare` needs the witness and
// the constraints but no challenge -- in the reference's order (commit, then zerocheck) it runs while the witness is uploaded
// (host_columns non-empty: chunked upload into multilinears[j].evals as in the streamed round) and committed;
// `zerocheck_univariate_finish` weights the prepared values by the eq-indicator of the challenges.  Values are identical to
// zerocheck_univariate_evals, which finish calls itself for shapes prepare does not cover.
struct PreparedUnivariateRound {
	std::vector<SumcheckMultilinear> multilinears;
	std::vector<const ExprEval *> compositions;
	std::vector<uint32_t> composition_degrees;
	uint32_t skip_rounds = 0, max_domain_size = 0;
	DevSlice store{};
	bool prepared = false, owns_store = true;
	void release(B200Backend &backend) {
		if (store.ptr && owns_store) backend.layer().dev_free(store);
		store = DevSlice{};
	}
};

inline PreparedUnivariateRound zerocheck_univariate_prepare(B200Backend &backend, const std::vector<const void *> &host_columns,
															const std::vector<SumcheckMultilinear> &multilinears,
															const std::vector<const ExprEval *> &compositions,
															const std::vector<uint32_t> &composition_degrees, uint32_t skip_rounds,
															uint32_t max_domain_size, uint32_t log_chunks = 3, const DevSlice *arena_store = nullptr) {
	if (multilinears.empty() || (!host_columns.empty() && host_columns.size() != multilinears.size()))
		throw InputValidation(1, "NumberOfVariablesMismatch: one host column per multilinear");
	const uint32_t n_vars = multilinears[0].n_vars;
	for (auto &m : multilinears)
		if (m.kind != SumcheckMultilinear::Transparent || m.n_vars != n_vars) throw InputValidation(1, "NumberOfVariablesMismatch");
	if (skip_rounds > n_vars) throw InputValidation(1, "TooManySkippedRounds");
	if (compositions.size() != composition_degrees.size()) throw InputValidation(1, "one degree per composition");
	uint32_t max_deg = 0;
	for (uint32_t d : composition_degrees) max_deg = std::max(max_deg, d);
	if ((uint64_t)max_domain_size < ((uint64_t)max_deg << skip_rounds)) throw InputValidation(1, "LagrangeDomainTooSmall");
	if (max_domain_size > 256) throw InputValidation(1, "DomainSizeTooLarge");
	PreparedUnivariateRound p;
	p.multilinears = multilinears, p.compositions = compositions, p.composition_degrees = composition_degrees;
	p.skip_rounds = skip_rounds, p.max_domain_size = max_domain_size;
	std::vector<b200_dev_ptr> ptrs;
	std::vector<uint32_t> levels;
	for (auto &m : multilinears) { ptrs.push_back(m.evals.ptr); levels.push_back(m.tower_level); }
	std::vector<const b200_expr *> exprs;
	for (auto *c : compositions) exprs.push_back(c->raw());
	B200Layer &l = backend.layer();
	const uint64_t n_store = b200_zerocheck_univariate_store_elems(n_vars, skip_rounds, composition_degrees.data(), (uint32_t)composition_degrees.size());
	// the value store comes from the caller's device arena when given (ComputeHolder's bump allocator, compute/src/alloc.rs),
	// else from a device allocation this object owns
	if (arena_store) {
		if (arena_store->n < n_store) throw InputValidation(1, "the store slice is smaller than b200_zerocheck_univariate_store_elems");
		p.store = arena_store->slice(0, n_store), p.owns_store = false;
	} else if (n_store)
		p.store = l.dev_alloc(n_store);
	uint32_t done = 0;
	try {
		l.check(b200_zerocheck_univariate_prepare(l.ctx(), host_columns.empty() ? nullptr : host_columns.data(), ptrs.data(), levels.data(), (uint32_t)ptrs.size(),
												  n_vars, skip_rounds, exprs.data(), composition_degrees.data(), (uint32_t)exprs.size(), max_domain_size, log_chunks,
												  p.store.ptr, n_store, &done));
	} catch (...) {
		p.release(backend);
		throw;
	}
	p.prepared = done != 0;
	if (!p.prepared) p.release(backend);
	return p;
}

inline ZerocheckUnivariateEvalsOutput zerocheck_univariate_finish(B200Backend &backend, const PreparedUnivariateRound &p,
																  const std::vector<F128> &zerocheck_challenges) {
	if (!p.prepared) return zerocheck_univariate_evals(backend, p.multilinears, p.compositions, p.composition_degrees, zerocheck_challenges, p.skip_rounds, p.max_domain_size);
	const uint32_t n_vars = p.multilinears[0].n_vars;
	if (zerocheck_challenges.size() != n_vars - p.skip_rounds) throw InputValidation(1, "IncorrectZerocheckChallengesLength");
	ZerocheckUnivariateEvalsOutput out;
	out.skip_rounds = p.skip_rounds, out.remaining_rounds = n_vars - p.skip_rounds, out.max_domain_size = p.max_domain_size;
	out.partial_eq_ind_evals = backend.tensor_product_full_query(zerocheck_challenges);
	std::vector<b200_dev_ptr> ptrs;
	std::vector<uint32_t> levels;
	for (auto &m : p.multilinears) { ptrs.push_back(m.evals.ptr); levels.push_back(m.tower_level); }
	std::vector<const b200_expr *> exprs;
	for (auto *c : p.compositions) exprs.push_back(c->raw());
	const uint32_t n_out = p.max_domain_size - (1u << p.skip_rounds);
	std::vector<F128> flat(std::max<size_t>(p.compositions.size() * n_out, 1));
	B200Layer &l = backend.layer();
	l.check(b200_zerocheck_univariate_finish(l.ctx(), ptrs.data(), levels.data(), (uint32_t)ptrs.size(), n_vars, p.skip_rounds, out.partial_eq_ind_evals.ptr,
											 out.partial_eq_ind_evals.n, exprs.data(), p.composition_degrees.data(), (uint32_t)exprs.size(), p.max_domain_size,
											 p.store.ptr, p.store.n, (uint64_t *)flat.data()));
	for (size_t c = 0; c < p.compositions.size(); c++) out.round_evals.emplace_back(flat.begin() + c * n_out, flat.begin() + (c + 1) * n_out);
	return out;
}

}  // namespace binius_b200
