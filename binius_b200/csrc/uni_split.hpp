// uni_split.hpp -- host-side planning for the univariate-skip fast kernel when a call's columns do not fit one
// CTA's shared memory: constraints are local (a chi constraint touches 4 columns), so the compositions are cut
// into contiguous ranges whose referenced columns fit, and each range is launched on its compacted column list
// with remapped monomials.  Pure C++ (no CUDA types) so that tests/cpp/uni_split_test.cpp can check it on the CPU.
#pragma once
#include <cstdint>
#include <set>
#include <vector>

namespace b200 {
namespace uni {

constexpr uint32_t SPLIT_MONO_NONE = 511;  // == MONO_NONE (univariate.cuh)
constexpr uint32_t SPLIT_CTAB = 5;         // == CTAB: first monomial, #quadratic, #linear, #general, points

struct MonoW {
	uint32_t x, y;  // layout of the kernel's uint2 monomial descriptor
};
struct SplitRange {
	uint32_t c0 = 0, c1 = 0;
	std::vector<uint32_t> cols;  // original column indices, ascending; local index = position
	std::vector<MonoW> mono;     // remapped descriptors of compositions [c0, c1)
	std::vector<uint32_t> ctab;  // SPLIT_CTAB words per composition, `first` relative to `mono`
};

// columns referenced by composition c; `cube` = SUBS * K (typed runs store column * cube)
inline void columns_of(const std::vector<MonoW> &mono, const std::vector<uint32_t> &ctab, uint32_t c, uint32_t cube, std::set<uint32_t> &cols) {
	const uint32_t *ct = &ctab[SPLIT_CTAB * c];
	for (uint32_t t = 0; t < ct[1] + ct[2] + ct[3]; t++) {
		const MonoW d = mono[ct[0] + t];
		if (t < ct[1]) {
			cols.insert(d.x / cube);
			cols.insert(d.y / cube);
		} else if (t < ct[1] + ct[2]) {
			cols.insert(d.x / cube);
		} else {
			if ((d.x & 511u) != SPLIT_MONO_NONE) cols.insert(d.x & 511u);
			if (((d.x >> 9) & 511u) != SPLIT_MONO_NONE) cols.insert((d.x >> 9) & 511u);
		}
	}
}

// fits(n_columns, n_compositions) -> bool.  Returns false when a single composition does not fit.
template <class Fits>
bool plan_split(const std::vector<MonoW> &mono, const std::vector<uint32_t> &ctab, uint32_t n_comp, uint32_t n_cols_total, uint32_t cube, Fits fits,
				std::vector<SplitRange> &out) {
	out.clear();
	for (uint32_t c0 = 0; c0 < n_comp;) {
		std::set<uint32_t> cols;
		uint32_t c1 = c0;
		while (c1 < n_comp) {
			std::set<uint32_t> t = cols;
			columns_of(mono, ctab, c1, cube, t);
			if (!fits((uint32_t)(t.empty() ? 1 : t.size()), c1 + 1 - c0)) break;
			cols.swap(t);
			c1++;
		}
		if (c1 == c0) return false;
		SplitRange R;
		R.c0 = c0, R.c1 = c1;
		R.cols.assign(cols.begin(), cols.end());
		std::vector<uint32_t> local(n_cols_total, 0);
		for (uint32_t i = 0; i < R.cols.size(); i++) local[R.cols[i]] = i;
		R.ctab.resize(SPLIT_CTAB * (size_t)(c1 - c0));
		for (uint32_t c = c0; c < c1; c++) {
			const uint32_t *ct = &ctab[SPLIT_CTAB * c];
			uint32_t *o = &R.ctab[SPLIT_CTAB * (c - c0)];
			o[0] = (uint32_t)R.mono.size(), o[1] = ct[1], o[2] = ct[2], o[3] = ct[3], o[4] = ct[4];
			for (uint32_t t = 0; t < ct[1] + ct[2] + ct[3]; t++) {
				MonoW d = mono[ct[0] + t];
				if (t < ct[1]) d = MonoW{local[d.x / cube] * cube, local[d.y / cube] * cube};
				else if (t < ct[1] + ct[2]) d = MonoW{local[d.x / cube] * cube, 0};
				else {
					uint32_t a = d.x & 511u, b = (d.x >> 9) & 511u;
					if (a != SPLIT_MONO_NONE) a = local[a];
					if (b != SPLIT_MONO_NONE) b = local[b];
					d = MonoW{a | (b << 9) | (d.x & ~0x3FFFFu), 0};
				}
				R.mono.push_back(d);
			}
		}
		out.push_back(std::move(R));
		c0 = c1;
	}
	return true;
}

}  // namespace uni
}  // namespace b200
