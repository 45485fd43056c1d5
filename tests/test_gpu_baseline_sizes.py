"""GPU parity tests AT THE SIZES bench.py measures (BASELINE.json configs #2, #3 and the keccak-shaped
calls): every benchmarked configuration is compared with the oracle bit for bit, not only through
size-independent properties.

  cfg#3  u32_add zerocheck: 5 multilinears of 18 variables, the two gadget compositions, all 18 rounds
  a2     bivariate round evaluations: m = 8 multilinears, 8 random index pairs, n = 20 (fused and traced)
  cfg#2  additive NTT over B32 at 2^24 coefficients: S1 (6,18,0, skip 1) on strided columns (each x is
         an independent 2^18 transform, crates/ntt/src/tests/reference.rs:197-214), S2 (0,24,0) whole
  f1     univariate-skip round at the reference's skip = 7 with 153 B1 columns and 75 chi-shaped
         constraints (m3/src/gadgets/hash/keccak/stacked.rs:340-366) over 2^16 rows
"""
import random

import numpy as np
import pytest

from test_gpu_hal import _same, u32_add_compositions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hal():
    import binius_b200

    layer = binius_b200.B200Layer(0)
    yield layer
    layer.close()


@pytest.mark.parametrize("tail", [False, True])
def test_cfg3_u32_add_zerocheck_n18_every_round(hal, oracle, tail):
    """tail=True: the whole sumcheck (all 18 rounds, from the first one with its own point set) runs in ONE persistent
    kernel on a co-resident grid (csrc/tail_grid.cuh); the multilinears are only readable after the last round."""
    from binius_b200.hal import B200Backend, EqIndEvaluator, FoldedMultilinear

    be = B200Backend(hal, sumcheck_tail=tail)
    n_vars = 18
    rng = random.Random(1803)
    comps = u32_add_compositions()
    mls_h = [oracle.rand_b128(1800 + t, 1 << n_vars) for t in range(5)]
    eq_pt = [rng.getrandbits(128) for _ in range(n_vars - 1)]
    eq_h = oracle.tensor_expand(oracle.to_arr([1] + [0] * ((1 << (n_vars - 1)) - 1)), 0, eq_pt)
    mls = [FoldedMultilinear(hal.to_device(m), 0) for m in mls_h]
    eq_d = be.tensor_product_full_query(eq_pt)
    assert _same(hal.to_host(eq_d), eq_h)
    for rnd in range(n_vars):
        nv = n_vars - rnd
        evs = [EqIndEvaluator(c, have_first_round_eval_1s=(rnd == 0)) for c in comps]
        got = be.sumcheck_compute_round_evals(nv, mls, evs, eq_d, [])
        exp = oracle.eq_ind_round_evals(mls_h, [len(m) for m in mls_h], [0] * 5, nv, eq_h, [c.steps for c in comps],
                                        [c.leading_term().steps for c in comps], [1, 2], [0, 0])
        for ev, g, e in zip(evs, got, exp):
            assert g == [e[k - 1] for k in ev.eval_point_indices()], f"round {rnd}"
        ch = rng.getrandbits(128)
        be.sumcheck_fold_multilinears(nv, mls, ch)
        mls_h = [oracle.fold_left_lerp_inplace(m, len(m), 0, nv, ch) for m in mls_h]
        assert (be._tail is not None) == (tail and nv > 1), "the persistent kernel must cover every round from the first"
        if not tail and rnd in (0, 1, 5, 12, n_vars - 1):  # full download of the folded multilinears on a few rounds
            for d, h in zip(mls, mls_h):
                assert d.evals.len() == len(h) and _same(hal.to_host(d.evals), h), f"fold of round {rnd}"
        if nv > 1:
            eq_d = be.fold_partial_eq_ind(nv - 1, eq_d)
            eq_h = oracle.fold_partial_eq_ind(eq_h)
    for d, h in zip(mls, mls_h):
        assert _same(hal.to_host(d.evals), h)


def test_bivariate_round_evals_m8_n20(hal, oracle):
    """v3::calculate_round_evals (bivariate_product.rs:303-408) at the size bench.py times: the fused entry point
    and the traced accumulate_kernels route must both equal the oracle."""
    from binius_b200.layer import calculate_round_evals

    n_vars, m = 20, 8
    rng = random.Random(2008)
    mls_h = [oracle.rand_b128(2000 + t, 1 << n_vars) for t in range(m)]
    mls = [hal.to_device(x) for x in mls_h]
    pairs = [(rng.randrange(m), rng.randrange(m)) for _ in range(8)]
    coeff = rng.getrandbits(128)
    exp = oracle.bivariate_round_evals(mls_h, n_vars, pairs, coeff)
    got = hal.execute(lambda ex: list(ex.bivariate_round_evals(mls, n_vars, pairs, coeff)))
    assert got == exp
    before = hal.launch_count()
    traced = calculate_round_evals(hal, n_vars, coeff, mls, pairs)
    assert traced == exp
    # the traced route must reach the tensor-core kernel: k_pair_tc + combine (+ the gmat memset is not a kernel)
    assert hal.launch_count() - before <= 3


def test_ntt_2pow24_against_oracle(hal, oracle):
    import binius_b200

    ntt = binius_b200.B200AdditiveNTT(hal, 5, 24)
    ontt = oracle.NTT(5, 24)
    rng = np.random.default_rng(24)
    n = 1 << 24
    a = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    # S1: log_x 6, log_y 18, skip 1 -- column x is the independent transform of a[x::64]
    S1 = binius_b200.NTTShape(6, 18, 0)
    f = a.copy()
    ntt.forward_transform(f, S1, 0, 0, 1)
    for x in (0, 17, 63):
        assert np.array_equal(f[x::64], ontt.forward(np.ascontiguousarray(a[x::64]), 5, 0, 18, 0, 0, 0, 1)), f"S1 forward, column {x}"
    g = a.copy()
    ntt.inverse_transform(g, S1, 0, 0, 1)
    for x in (5, 40):
        assert np.array_equal(g[x::64], ontt.inverse(np.ascontiguousarray(a[x::64]), 5, 0, 18, 0, 0, 0, 1)), f"S1 inverse, column {x}"
    # S2: one transform of 2^24 points, forward and inverse, the whole vector
    S2 = binius_b200.NTTShape(0, 24, 0)
    f = a.copy()
    ntt.forward_transform(f, S2)
    assert np.array_equal(f, ontt.forward(a, 5, 0, 24, 0))
    g = a.copy()
    ntt.inverse_transform(g, S2)
    assert np.array_equal(g, ontt.inverse(a, 5, 0, 24, 0))
    # RS-encode shape of the prover (reed_solomon.rs:143-157): a coset of a larger domain
    S1c = binius_b200.NTTShape(6, 17, 0)
    f = a[: 1 << 23].copy()
    ntt.forward_transform(f, S1c, 1, 1, 0)
    for x in (3, 62):
        assert np.array_equal(f[x::64], ontt.forward(np.ascontiguousarray(a[: 1 << 23][x::64]), 5, 0, 17, 0, 1, 1, 0)), f"coset, column {x}"


def keccak_chi_compositions():
    """75 constraints out - (b0 + (b1 - 1) * b2) over 153 columns: 75 state_out, 75 b, round constant, 2 spare
    (stacked.rs:318-366: per batch 25 state_out + 25 b + round_const; three batches share the table)."""
    from binius_b200 import ArithCircuit as A

    v = [A.var(i) for i in range(153)]
    comps = []
    for batch in range(3):
        for xy in range(25):
            x, y = xy % 5, xy // 5
            b = [v[75 + 25 * batch + ((x + k) % 5) + 5 * y] for k in range(3)]
            comps.append(v[25 * batch + xy] - (b[0] + (b[1] - A.one()) * b[2]))
    return comps


def test_univariate_skip7_153_columns_2pow16_rows(hal, oracle):
    """The reference picks skip_rounds = 7 for degree-2 constraints over B8 (constraint_system/verify.rs:271-294).
    Checker: the threaded table-driven CPU arm (oracle/cpu_univariate.c), itself compared with the definition at
    skip 3..7 in tests/test_oracle_univariate.py."""
    from binius_b200.hal import B200Backend, TransparentMultilinear, zerocheck_univariate_evals

    be = B200Backend(hal)
    n_vars, skip, m = 16, 7, 153
    rng = random.Random(153)
    cols = [oracle.rand_b128(1530 + j, (1 << n_vars) // 128) for j in range(m)]
    comps = keccak_chi_compositions()
    ch = [rng.getrandbits(128) for _ in range(n_vars - skip)]
    mls = [TransparentMultilinear(hal.to_device(c), 0, n_vars) for c in cols]
    eq = oracle.tensor_expand(oracle.to_arr([1] + [0] * ((1 << len(ch)) - 1)), 0, ch)
    exp, _ = oracle.cpu_univariate_b1(cols, n_vars, skip, eq, [list(c.steps) for c in comps], 1 << skip)
    for generic in (0, 1):
        hal.set_tuning("uni_generic", generic)
        try:
            out = zerocheck_univariate_evals(be, mls, comps, ch, skip, 256)
        finally:
            hal.set_tuning("uni_generic", 0)
        assert np.array_equal(hal.to_host(out.partial_eq_ind_evals), eq)
        assert out.round_evals == exp, "generic kernel" if generic else "split fast path"


def test_univariate_skip7_linear_monomial_route_2pow20_rows(hal, oracle):
    """From 2^12 sub-cubes of 128 rows on, coefficient-1 linear monomials leave k_uni_b8: their sums are
    evaluate_partial_high of the column by the eq-indicator (tensor-core outer product) folded with the Lagrange
    coefficients (k_uni_linear).  The keccak shape plus compositions with a constant term, a scaled linear term and a
    purely linear composition, against the CPU arm and against the kernel with the route switched off."""
    from binius_b200 import ArithCircuit as A
    from binius_b200.hal import B200Backend, TransparentMultilinear, zerocheck_univariate_evals

    be = B200Backend(hal)
    n_vars, skip, m = 20, 7, 153
    rng = random.Random(2020)
    cols = [oracle.rand_b128(2530 + j, (1 << n_vars) // 128) for j in range(m)]
    v = [A.var(i) for i in range(m)]
    comps = keccak_chi_compositions() + [v[150] * v[151] + v[152] + A.constant(0x53), v[3] * v[77] + A.constant(0x0B) * v[151] + v[0],
                                         v[1] + v[2] + v[152]]
    ch = [rng.getrandbits(128) for _ in range(n_vars - skip)]
    mls = [TransparentMultilinear(hal.to_device(c), 0, n_vars) for c in cols]
    eq = oracle.tensor_expand(oracle.to_arr([1] + [0] * ((1 << len(ch)) - 1)), 0, ch)
    exp, _ = oracle.cpu_univariate_b1(cols, n_vars, skip, eq, [list(c.steps) for c in comps[:-1]], 1 << skip)
    got = {}
    for lin in (1, 0):
        hal.set_tuning("uni_linear", lin)
        try:
            before = hal.launch_count()
            got[lin] = zerocheck_univariate_evals(be, mls, comps, ch, skip, 256).round_evals
            launches = hal.launch_count() - before
        finally:
            hal.set_tuning("uni_linear", 1)
        if lin:
            routed = launches
    assert got[1] == got[0], "linear-monomial route vs all monomials in the kernel"
    assert got[1][:-1] == exp, "vs the CPU arm"
    assert routed > launches, "the route adds the outer-product, combine and k_uni_linear launches"
