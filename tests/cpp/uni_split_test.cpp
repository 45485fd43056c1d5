// CPU unit test of the composition-range split planner (binius_b200/csrc/uni_split.hpp): every composition in
// exactly one range, each range within the budget, remapped monomials referring to the original columns.
#include <cstdio>
#include <cstdlib>

#include "../../binius_b200/csrc/uni_split.hpp"

using namespace b200::uni;
#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static uint64_t st = 12345;
static uint32_t rnd(uint32_t n) {
	st = st * 6364136223846793005ull + 1442695040888963407ull;
	return (uint32_t)((st >> 33) % n);
}

int main() {
	const uint32_t cube = 8 * 128;
	for (int trial = 0; trial < 200; trial++) {
		const uint32_t m = 20 + rnd(200), n_comp = 1 + rnd(100), budget = 8 + rnd(60), max_comps = 1 + rnd(40);
		std::vector<MonoW> mono;
		std::vector<uint32_t> ctab(SPLIT_CTAB * n_comp);
		for (uint32_t c = 0; c < n_comp; c++) {
			uint32_t *ct = &ctab[SPLIT_CTAB * c];
			ct[0] = (uint32_t)mono.size(), ct[1] = rnd(3), ct[2] = rnd(3), ct[3] = rnd(3), ct[4] = 128;
			const uint32_t base = rnd(m);  // local constraints: columns near `base`
			auto col = [&]() { return (base + rnd(6)) % m; };
			for (uint32_t t = 0; t < ct[1]; t++) mono.push_back(MonoW{col() * cube, col() * cube});
			for (uint32_t t = 0; t < ct[2]; t++) mono.push_back(MonoW{col() * cube, 0});
			for (uint32_t t = 0; t < ct[3]; t++) {
				const uint32_t kind = rnd(3), a = kind > 0 ? col() : SPLIT_MONO_NONE, b = kind > 1 ? col() : SPLIT_MONO_NONE;
				mono.push_back(MonoW{a | (b << 9) | ((1 + rnd(255)) << 18), 0});
			}
		}
		std::vector<SplitRange> ranges;
		const bool ok = plan_split(mono, ctab, n_comp, m, cube, [&](uint32_t ncols, uint32_t ncomp) { return ncols <= budget && ncomp <= max_comps; }, ranges);
		CHECK(ok);  // a single composition touches at most 12 columns > ... budget >= 8 may fail: handled below
		uint32_t next = 0;
		for (const SplitRange &R : ranges) {
			CHECK(R.c0 == next && R.c1 > R.c0 && R.c1 <= n_comp);
			next = R.c1;
			CHECK(R.cols.size() <= budget && R.c1 - R.c0 <= max_comps);
			for (size_t i = 1; i < R.cols.size(); i++) CHECK(R.cols[i - 1] < R.cols[i] && R.cols[i] < m);
			CHECK(R.ctab.size() == SPLIT_CTAB * (size_t)(R.c1 - R.c0));
			uint32_t running = 0;
			for (uint32_t c = R.c0; c < R.c1; c++) {
				const uint32_t *ct = &ctab[SPLIT_CTAB * c], *o = &R.ctab[SPLIT_CTAB * (c - R.c0)];
				CHECK(o[0] == running && o[1] == ct[1] && o[2] == ct[2] && o[3] == ct[3] && o[4] == ct[4]);
				for (uint32_t t = 0; t < ct[1] + ct[2] + ct[3]; t++) {
					const MonoW d = mono[ct[0] + t], e = R.mono[o[0] + t];
					if (t < ct[1]) {
						CHECK(e.x % cube == 0 && e.y % cube == 0 && e.x / cube < R.cols.size() && e.y / cube < R.cols.size());
						CHECK(R.cols[e.x / cube] == d.x / cube && R.cols[e.y / cube] == d.y / cube);
					} else if (t < ct[1] + ct[2]) {
						CHECK(e.x % cube == 0 && e.x / cube < R.cols.size() && R.cols[e.x / cube] == d.x / cube && e.y == 0);
					} else {
						const uint32_t a = d.x & 511u, b = (d.x >> 9) & 511u, la = e.x & 511u, lb = (e.x >> 9) & 511u;
						CHECK((e.x >> 18) == (d.x >> 18));
						CHECK(a == SPLIT_MONO_NONE ? la == SPLIT_MONO_NONE : (la < R.cols.size() && R.cols[la] == a));
						CHECK(b == SPLIT_MONO_NONE ? lb == SPLIT_MONO_NONE : (lb < R.cols.size() && R.cols[lb] == b));
					}
				}
				running += ct[1] + ct[2] + ct[3];
			}
			CHECK(R.mono.size() == running);
		}
		CHECK(next == n_comp);
	}
	// a composition that cannot fit on its own is reported, not split
	{
		std::vector<MonoW> mono = {MonoW{0 * cube, 1 * cube}, MonoW{2 * cube, 3 * cube}};
		std::vector<uint32_t> ctab = {0, 2, 0, 0, 128};
		std::vector<SplitRange> ranges;
		CHECK(!plan_split(mono, ctab, 1, 4, cube, [](uint32_t ncols, uint32_t) { return ncols <= 3; }, ranges));
	}
	printf("uni split ok\n");
	return 0;
}
