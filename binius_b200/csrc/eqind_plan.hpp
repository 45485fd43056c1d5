// eqind_plan.hpp -- host-side symbolic expansion of ArithCircuit compositions into monomials, used to
// lower eq-ind (zerocheck) round evaluations of degree <= 2 compositions to tensor-core inner
// products:
//     sum_i E[i] * C(P(i)) = sum_k coef_k * sum_i E[i] * prod_{v in S_k} P_v(i)
//       |S_k| = 0 :  coef * sum_i E[i]                        -> job (E, ones)
//       |S_k| = 1 :  coef * sum_i E[i] * P_x(i)               -> job (E, P_x)
//       |S_k| = 2 :  coef * sum_i (E[i] * P_x(i)) * P_y(i)    -> one elementwise product w = E . P_x per
//                                                               distinct scaled factor, then job (w, P_y)
// (reference semantics: hal/src/sumcheck_round_calculation.rs:222-297, eq_ind.rs:670-702; the
// ArithCircuit steps are math/src/arith_expr.rs:200-206).  Independent of oracle/.
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <vector>

#include "../../include/binius_b200.h"
#include "host_field.hpp"

namespace b200 {
namespace plan {

using hostf::u128;
typedef std::vector<uint32_t> Mono;  // sorted variable indices, with multiplicity
typedef std::map<Mono, u128> Poly;

inline void add_term(Poly &p, const Mono &m, u128 c) {
	if (!c) return;
	auto it = p.find(m);
	if (it == p.end()) p[m] = c;
	else {
		it->second ^= c;
		if (!it->second) p.erase(it);
	}
}

inline bool mul(const Poly &a, const Poly &b, uint32_t max_degree, size_t max_terms, Poly &out) {
	out.clear();
	for (auto &ta : a)
		for (auto &tb : b) {
			Mono m(ta.first);
			m.insert(m.end(), tb.first.begin(), tb.first.end());
			if (m.size() > max_degree) return false;
			std::sort(m.begin(), m.end());
			add_term(out, m, hostf::mul128(ta.second, tb.second));
			if (out.size() > max_terms) return false;
		}
	return true;
}

// false when the polynomial has degree > max_degree or more than max_terms monomials
inline bool expand(const b200_expr_step *steps, uint32_t n_steps, uint32_t max_degree, size_t max_terms, Poly &out) {
	std::vector<Poly> v(n_steps);
	for (uint32_t s = 0; s < n_steps; s++) {
		const b200_expr_step &st = steps[s];
		switch (st.op) {
		case 0:
			v[s] = v[st.l];
			for (auto &t : v[st.r]) add_term(v[s], t.first, t.second);
			if (v[s].size() > max_terms) return false;
			break;
		case 1:
			if (!mul(v[st.l], v[st.r], max_degree, max_terms, v[s])) return false;
			break;
		case 2: {
			Poly acc;
			acc[Mono()] = 1;
			if (st.r > max_degree && !v[st.l].empty()) {
				// only constants may be raised to large powers
				for (auto &t : v[st.l])
					if (!t.first.empty()) return false;
			}
			for (uint64_t e = 0; e < st.r; e++) {
				Poly nx;
				if (!mul(acc, v[st.l], max_degree, max_terms, nx)) return false;
				acc.swap(nx);
				if (acc.empty()) break;
				if (e > 1024) return false;
			}
			v[s] = acc;
			break;
		}
		case 3: {
			u128 c = ((u128)st.c_hi << 64) | st.c_lo;
			if (c) v[s][Mono()] = c;
			break;
		}
		default: v[s][Mono{st.l}] = 1; break;
		}
	}
	out = n_steps ? v[n_steps - 1] : Poly();
	return true;
}

}  // namespace plan
}  // namespace b200
