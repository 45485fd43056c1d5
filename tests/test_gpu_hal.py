"""GPU parity tests of the old-HAL (ComputationBackend) hot path: eq-ind round evaluations, fold of
Folded multilinears with constant suffixes, eq-ind expansion -- CUDA path vs oracle, round by round.
Workload shape = BASELINE config #3 scaled down (u32_add constraints, m3/src/gadgets/add.rs:71-76)."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hal():
    import binius_b200

    layer = binius_b200.B200Layer(0)
    yield layer
    layer.close()


def _same(a, b):
    return np.array_equal(np.asarray(a, dtype=np.uint64).reshape(-1, 2), np.asarray(b, dtype=np.uint64).reshape(-1, 2))


def u32_add_compositions():
    from binius_b200 import ArithCircuit as A

    x, y, cin, cout, z = (A.var(i) for i in range(5))
    c1 = (x + cin) * (y + cin) + cin - cout  # degree 2
    c2 = x + y + cin - z  # degree 1
    return [c1, c2]


def test_tensor_product_full_query(hal, oracle):
    from binius_b200.hal import B200Backend

    be = B200Backend(hal)
    rng = random.Random(1)
    for k in (0, 1, 4, 11, 13):
        q = [rng.getrandbits(128) for _ in range(k)]
        out = be.tensor_product_full_query(q)
        exp = oracle.tensor_expand(oracle.to_arr([1] + [0] * ((1 << k) - 1)), 0, q)
        assert _same(hal.to_host(out), exp)


@pytest.mark.parametrize("n_vars", [1, 5, 11, 14])
def test_zerocheck_rounds_match_oracle(hal, oracle, n_vars):
    """All rounds of an eq-ind sumcheck: round evals (at 1 and infinity, plus an extra finite domain
    point for the degree-3 variant), fold of every multilinear, halving of the eq-ind table."""
    from binius_b200 import ArithCircuit as A
    from binius_b200.hal import B200Backend, EqIndEvaluator, FoldedMultilinear

    be = B200Backend(hal)
    rng = random.Random(n_vars)
    comps = u32_add_compositions() + [A.var(0) * A.var(1) * A.var(4) + A.var(2)]  # a degree-3 one -> finite point
    mls_h = [oracle.rand_b128(200 + t, 1 << n_vars) for t in range(5)]
    eq_pt = [rng.getrandbits(128) for _ in range(n_vars - 1)]
    eq_h = oracle.tensor_expand(oracle.to_arr([1] + [0] * ((1 << (n_vars - 1)) - 1)), 0, eq_pt)
    mls = [FoldedMultilinear(hal.to_device(m), 0) for m in mls_h]
    eq_d = be.tensor_product_full_query(eq_pt)
    assert _same(hal.to_host(eq_d), eq_h)
    finite = [oracle.mul(0x2, 0x2)]  # domain point index 3
    for rnd in range(n_vars):
        nv = n_vars - rnd
        evs = [EqIndEvaluator(c, have_first_round_eval_1s=(rnd == 0 and i == 0)) for i, c in enumerate(comps)]
        got = be.sumcheck_compute_round_evals(nv, mls, evs, eq_d, finite)
        codes = [1, 2, 3]
        exp_all = oracle.eq_ind_round_evals(mls_h, [len(m) for m in mls_h], [0] * 5, nv, eq_h,
                                            [c.steps for c in comps], [c.leading_term().steps for c in comps], codes, [0, 0, finite[0]])
        for ev, g, e in zip(evs, got, exp_all):
            assert g == [e[k - 1] for k in ev.eval_point_indices()]
        assert len(got[1]) == 1 and len(got[2]) == 3 and len(got[0]) == (1 if rnd == 0 else 2)
        ch = rng.getrandbits(128)
        be.sumcheck_fold_multilinears(nv, mls, ch)
        mls_h = [oracle.fold_left_lerp_inplace(m, len(m), 0, nv, ch) for m in mls_h]
        for d, h in zip(mls, mls_h):
            assert d.evals.len() == len(h) and _same(hal.to_host(d.evals), h)
        if nv > 1:
            eq_d = be.fold_partial_eq_ind(nv - 1, eq_d)
            eq_h = oracle.fold_partial_eq_ind(eq_h)
            assert _same(hal.to_host(eq_d), eq_h)
    assert all(m.evals.len() == 1 for m in mls)


def test_zerocheck_degree2_monomial_plan(hal, oracle):
    """Degree <= 2 compositions at the points 1 / infinity on large rounds take the monomial plan
    (shared tensor-core inner products, eqind_plan.hpp); smaller rounds fall back to the interpreter.
    Covers shared variables, a square, a constant term, a scaled monomial and an identically-zero
    composition (x*y + y*x); shape = keccak chi constraints (m3/src/gadgets/hash/keccak/stacked.rs)."""
    from binius_b200 import ArithCircuit as A
    from binius_b200.hal import B200Backend, EqIndEvaluator, FoldedMultilinear

    be = B200Backend(hal)
    n_vars, m = 14, 9
    rng = random.Random(77)
    v = [A.var(i) for i in range(m)]
    comps = [v[c] - (v[4 + c % 5] + (v[4 + (c + 1) % 5] - A.one()) * v[4 + (c + 2) % 5]) for c in range(4)]
    comps += u32_add_compositions()
    comps += [v[3] * v[3] + A.constant(rng.getrandbits(128)) * v[2] * v[8] + A.constant(rng.getrandbits(128)),
              v[0] * v[1] + v[1] * v[0], v[5].pow(2) + v[6]]
    mls_h = [oracle.rand_b128(500 + t, 1 << n_vars) for t in range(m)]
    eq_pt = [rng.getrandbits(128) for _ in range(n_vars - 1)]
    eq_d = be.tensor_product_full_query(eq_pt)
    eq_h = hal.to_host(eq_d)
    mls = [FoldedMultilinear(hal.to_device(x), 0) for x in mls_h]
    for rnd in range(4):
        nv = n_vars - rnd
        evs = [EqIndEvaluator(c, have_first_round_eval_1s=(rnd == 0)) for c in comps]
        got = be.sumcheck_compute_round_evals(nv, mls, evs, eq_d, [])
        exp = oracle.eq_ind_round_evals(mls_h, [len(x) for x in mls_h], [0] * m, nv, eq_h, [c.steps for c in comps],
                                        [c.leading_term().steps for c in comps], [1, 2], [0, 0])
        for ev, g, e in zip(evs, got, exp):
            assert g == [e[k - 1] for k in ev.eval_point_indices()]
        ch = rng.getrandbits(128)
        be.sumcheck_fold_multilinears(nv, mls, ch)
        mls_h = [oracle.fold_left_lerp_inplace(x, len(x), 0, nv, ch) for x in mls_h]
        eq_d = be.fold_partial_eq_ind(nv - 1, eq_d)
        eq_h = oracle.fold_partial_eq_ind(eq_h)


def test_truncated_multilinears_with_const_suffix(hal, oracle):
    """Folded multilinears store only a non-constant prefix (fold.rs:648-696, round calc :573-600)."""
    from binius_b200.hal import B200Backend, EqIndEvaluator, FoldedMultilinear

    be = B200Backend(hal)
    n_vars = 8
    rng = random.Random(9)
    prefixes = [256, 200, 129, 128, 77]
    suffixes = [0, rng.getrandbits(128), 1, rng.getrandbits(128), rng.getrandbits(128)]
    mls_h = [oracle.rand_b128(300 + t, p) for t, p in enumerate(prefixes)]
    mls = [FoldedMultilinear(hal.to_device(m), s) for m, s in zip(mls_h, suffixes)]
    comps = u32_add_compositions()
    eq_pt = [rng.getrandbits(128) for _ in range(n_vars - 1)]
    eq_d = be.tensor_product_full_query(eq_pt)
    eq_h = hal.to_host(eq_d)
    for rnd in range(4):
        nv = n_vars - rnd
        evs = [EqIndEvaluator(c) for c in comps]
        got = be.sumcheck_compute_round_evals(nv, mls, evs, eq_d, [])
        exp = oracle.eq_ind_round_evals(mls_h, [len(m) for m in mls_h], suffixes, nv, eq_h, [c.steps for c in comps],
                                        [c.leading_term().steps for c in comps], [1, 2], [0, 0])
        assert got[0] == exp[0] and got[1] == exp[1][:1]
        ch = rng.getrandbits(128)
        be.sumcheck_fold_multilinears(nv, mls, ch)
        mls_h = [oracle.fold_left_lerp_inplace(m, len(m), s, nv, ch) for m, s in zip(mls_h, suffixes)]
        for d, h in zip(mls, mls_h):
            assert d.evals.len() == len(h) and _same(hal.to_host(d.evals), h)
        eq_d = be.fold_partial_eq_ind(nv - 1, eq_d)
        eq_h = oracle.fold_partial_eq_ind(eq_h)
    import binius_b200

    with pytest.raises(binius_b200.InputValidation):
        be.sumcheck_compute_round_evals(4, mls, [EqIndEvaluator(comps[0])], eq_d, [1, 2])


def test_evaluate_partial_high(hal, oracle):
    from binius_b200.hal import B200Backend

    be = B200Backend(hal)
    n, k = 10, 3
    ml = oracle.rand_b128(400, 1 << n)
    q = [5, 1 << 90, 0xABCDEF]
    qe = be.tensor_product_full_query(q)
    out = be.evaluate_partial_high(hal.to_device(ml), qe)
    exp = oracle.fold_left(ml, 7, hal.to_host(qe), 1 << (n - k))
    assert _same(hal.to_host(out), exp)
    # successive single-variable folds with the same challenges (reversed order: highest first) agree
    cur = ml
    for i, c in enumerate(reversed(q)):
        cur = oracle.fold_left_lerp_inplace(cur, len(cur), 0, n - i, c)
    assert _same(cur, exp)


@pytest.mark.parametrize("n_vars", [1, 6, 13])
@pytest.mark.parametrize("weighted", [True, False])
def test_low_to_high_rounds_match_oracle(hal, oracle, n_vars, weighted):
    """EvaluationOrder::LowToHigh (LowToHighAccess, fold_right_lerp, fold_partial_eq_ind low-to-high) with
    the eq-ind evaluator and with the regular (unweighted) evaluator, all rounds, truncated inputs."""
    from binius_b200 import ArithCircuit as A
    from binius_b200.hal import (B200Backend, EqIndEvaluator, EvaluationOrder, FoldedMultilinear,
                                 RegularSumcheckEvaluator)

    be = B200Backend(hal)
    rng = random.Random(100 + n_vars)
    comps = u32_add_compositions() + [A.var(0) * A.var(1) * A.var(4) + A.var(2)]
    full = 1 << n_vars
    prefixes = [full, full, max(full - 3, 1), max(full // 2 + 1, 1), max(full // 3, 1)]
    suffixes = [0, 0, rng.getrandbits(128), 1, rng.getrandbits(128)]
    mls_h = [oracle.rand_b128(900 + t, p) for t, p in enumerate(prefixes)]
    mls = [FoldedMultilinear(hal.to_device(m), s) for m, s in zip(mls_h, suffixes)]
    eq_pt = [rng.getrandbits(128) for _ in range(n_vars - 1)]
    eq_d = be.tensor_product_full_query(eq_pt) if weighted else None
    eq_h = hal.to_host(eq_d) if weighted else None
    finite = [oracle.mul(0x2, 0x2)]
    Ev = EqIndEvaluator if weighted else RegularSumcheckEvaluator
    for rnd in range(n_vars):
        nv = n_vars - rnd
        evs = [Ev(c) for c in comps]
        got = be.sumcheck_compute_round_evals(nv, mls, evs, eq_d, finite, evaluation_order=EvaluationOrder.LowToHigh)
        exp = oracle.sumcheck_round_evals(0, mls_h, [len(m) for m in mls_h], suffixes, nv, eq_h, [c.steps for c in comps],
                                          [c.leading_term().steps for c in comps], [1, 2, 3], [0, 0, finite[0]])
        for ev, g, e in zip(evs, got, exp):
            assert g == [e[k - 1] for k in ev.eval_point_indices()]
        ch = rng.getrandbits(128)
        assert be.sumcheck_fold_multilinears(nv, mls, ch, evaluation_order=EvaluationOrder.LowToHigh) is False
        mls_h = [oracle.fold_right_lerp(m, s, ch) for m, s in zip(mls_h, suffixes)]
        for d, h in zip(mls, mls_h):
            assert d.evals.len() == len(h) and _same(hal.to_host(d.evals), h)
        if weighted and nv > 1:
            eq_d = be.fold_partial_eq_ind(nv - 1, eq_d, EvaluationOrder.LowToHigh)
            eq_h = oracle.fold_partial_eq_ind_low_to_high(eq_h)
            assert _same(hal.to_host(eq_d), eq_h)
    assert all(m.evals.len() == 1 for m in mls)


def test_low_to_high_fold_rejects_overlap(hal, oracle):
    import ctypes as C

    import binius_b200

    d = hal.to_device(oracle.rand_b128(1, 64))
    ptrs = (C.c_void_p * 1)(d.ptr)
    lens = (C.c_uint64 * 1)(64)
    z = (C.c_uint64 * 2)(5, 0)
    sfx = (C.c_uint64 * 2)(0, 0)
    with pytest.raises(binius_b200.InputValidation):
        hal._check(hal._lib.b200_fold_multilinears_low_to_high(hal._ctx, ptrs, ptrs, 1, 6, lens, sfx, z, None))


def _pack_subfield(oracle, scalars, lvl):
    """2^(7-lvl) sub-field scalars of 2^lvl bits per B128 word, low limb first (memory.rs:257-281)."""
    per = 1 << (7 - lvl)
    words = []
    for w in range(len(scalars) // per):
        v = 0
        for j in range(per):
            v |= scalars[w * per + j] << (j << lvl)
        words.append(v)
    return oracle.to_arr(words)


@pytest.mark.parametrize("order_name", ["HighToLow", "LowToHigh"])
@pytest.mark.parametrize("lvl", [0, 3, 5])
def test_transparent_multilinears_switchover(hal, oracle, order_name, lvl):
    """SumcheckMultilinear::Transparent (sub-field multilinears, switchover rounds 0/1/2) next to a Folded
    one: every round's evaluations and the post-switchover folded values equal those of the same
    sumcheck run on the B128 embeddings (sumcheck_folding.rs:37-241, prover_state.rs:138-188)."""
    from binius_b200.hal import (B200Backend, EvaluationOrder, FoldedMultilinear, RegularSumcheckEvaluator,
                                 TransparentMultilinear)

    order = EvaluationOrder[order_name]
    be = B200Backend(hal)
    n_vars = 9
    rng = random.Random(lvl * 10 + int(order))
    bits = 1 << lvl
    scal = [[rng.getrandbits(bits) for _ in range(1 << n_vars)] for _ in range(3)]
    emb_h = [oracle.to_arr(s) for s in scal] + [oracle.rand_b128(33, 1 << n_vars)]
    mls = [TransparentMultilinear(hal.to_device(_pack_subfield(oracle, s, lvl)), lvl, n_vars, switchover_round=t)
           for t, s in enumerate(scal)] + [FoldedMultilinear(hal.to_device(emb_h[3]), 0)]
    from binius_b200 import ArithCircuit as A

    comps = [A.var(0) * A.var(1) + A.var(2) * A.var(3), A.var(0) * A.var(3) * A.var(2) + A.var(1)]
    finite = [oracle.mul(0x2, 0x2)]
    challenges = []
    for rnd in range(5):
        nv = n_vars - rnd
        # tensor query = expansion of the challenges so far (HighToLow: newest first, prover_state.rs:161-171)
        coords = challenges if order == EvaluationOrder.LowToHigh else list(reversed(challenges))
        tq = be.tensor_product_full_query(coords) if any(isinstance(x, TransparentMultilinear) for x in mls) else None
        evs = [RegularSumcheckEvaluator(c) for c in comps]
        got = be.sumcheck_compute_round_evals(nv, mls, evs, None, finite, evaluation_order=order, tensor_query=tq)
        exp = oracle.sumcheck_round_evals(int(order), emb_h, [len(x) for x in emb_h], [0] * 4, nv, None, [c.steps for c in comps],
                                          [c.leading_term().steps for c in comps], [1, 2, 3], [0, 0, finite[0]])
        for ev, g, e in zip(evs, got, exp):
            assert g == [e[k - 1] for k in ev.eval_point_indices()]
        ch = rng.getrandbits(128)
        challenges.append(ch)
        coords = challenges if order == EvaluationOrder.LowToHigh else list(reversed(challenges))
        tq = be.tensor_product_full_query(coords)
        left = be.sumcheck_fold_multilinears(nv, mls, ch, tq, evaluation_order=order)
        assert left == (rnd < 2)
        if order == EvaluationOrder.HighToLow:
            emb_h = [oracle.fold_left_lerp_inplace(x, len(x), 0, nv, ch) for x in emb_h]
        else:
            emb_h = [oracle.fold_right_lerp(x, 0, ch) for x in emb_h]
        for t, (d, h) in enumerate(zip(mls, emb_h)):
            assert isinstance(d, FoldedMultilinear) == (t <= rnd or t == 3)
            if isinstance(d, FoldedMultilinear):
                assert _same(hal.to_host(d.evals), h)


def test_regular_evaluator_monomial_plan(hal, oracle):
    """RegularSumcheckEvaluator (no eq-indicator weighting) on large HighToLow rounds takes the monomial
    plan with the all-ones vector in place of E: shared variables, a square, constants, a linear and an
    identically-zero composition; against the brute-force oracle, a few rounds."""
    from binius_b200 import ArithCircuit as A
    from binius_b200.hal import B200Backend, FoldedMultilinear, RegularSumcheckEvaluator

    be = B200Backend(hal)
    n_vars, m = 14, 6
    rng = random.Random(123)
    v = [A.var(i) for i in range(m)]
    comps = [v[0] * v[1] + v[2], v[3] * v[3] + A.constant(rng.getrandbits(128)) * v[4] * v[5] + A.constant(rng.getrandbits(128)),
             v[0] + v[5] + A.one(), v[1] * v[2] + v[2] * v[1], (v[0] + v[1]) * (v[2] + A.one())]
    mls_h = [oracle.rand_b128(1500 + t, 1 << n_vars) for t in range(m)]
    mls = [FoldedMultilinear(hal.to_device(x), 0) for x in mls_h]
    for rnd in range(3):
        nv = n_vars - rnd
        evs = [RegularSumcheckEvaluator(c) for c in comps]
        got = be.sumcheck_compute_round_evals(nv, mls, evs, None, [])
        exp = oracle.sumcheck_round_evals(1, mls_h, [len(x) for x in mls_h], [0] * m, nv, None, [c.steps for c in comps],
                                          [c.leading_term().steps for c in comps], [1, 2], [0, 0])
        for ev, g, e in zip(evs, got, exp):
            assert g == [e[k - 1] for k in ev.eval_point_indices()]
        ch = rng.getrandbits(128)
        be.sumcheck_fold_multilinears(nv, mls, ch)
        mls_h = [oracle.fold_left_lerp_inplace(x, len(x), 0, nv, ch) for x in mls_h]


@pytest.mark.parametrize("n_vars,with_cubic", [(1, False), (6, True), (13, False), (16, True)])
def test_persistent_sumcheck_tail(hal, oracle, n_vars, with_cubic):
    """The persistent tail kernel (b200_sumcheck_tail_*): once a round is small enough, all remaining rounds run in one
    kernel that talks to the host through mapped mailboxes.  Every round's values and the final folds must equal the
    oracle's; rounds above the threshold (n_vars = 16) still take the per-call path.  Includes a degree-3 composition
    (finite evaluation point 3) next to the u32_add ones."""
    from binius_b200 import ArithCircuit as A
    from binius_b200.hal import B200Backend, EqIndEvaluator, FoldedMultilinear

    be = B200Backend(hal, sumcheck_tail=True)
    be.tail_threshold = 1 << 14  # exercise the tail on rounds of up to 2^14 (composition, point, index) triples
    rng = random.Random(4000 + n_vars)
    comps = u32_add_compositions() + ([A.var(0) * A.var(1) * A.var(4) + A.var(2)] if with_cubic else [])
    mls_h = [oracle.rand_b128(4100 + t, 1 << n_vars) for t in range(5)]
    eq_pt = [rng.getrandbits(128) for _ in range(n_vars - 1)]
    eq_h = oracle.tensor_expand(oracle.to_arr([1] + [0] * ((1 << max(n_vars - 1, 0)) - 1)), 0, eq_pt)
    mls = [FoldedMultilinear(hal.to_device(m), 0) for m in mls_h]
    eq_d = be.tensor_product_full_query(eq_pt)
    finite = [oracle.mul(0x2, 0x2)] if with_cubic else []
    used_tail = False
    for rnd in range(n_vars):
        nv = n_vars - rnd
        evs = [EqIndEvaluator(c, have_first_round_eval_1s=(rnd == 0)) for c in comps]
        got = be.sumcheck_compute_round_evals(nv, mls, evs, eq_d, finite)
        used_tail = used_tail or be._tail is not None
        exp_all = oracle.eq_ind_round_evals(mls_h, [len(m) for m in mls_h], [0] * 5, nv, eq_h, [c.steps for c in comps],
                                            [c.leading_term().steps for c in comps], [1, 2, 3], [0, 0, finite[0] if finite else 0])
        for ev, g, e in zip(evs, got, exp_all):
            assert g == [e[k - 1] for k in ev.eval_point_indices()], f"round {rnd}"
        ch = rng.getrandbits(128)
        be.sumcheck_fold_multilinears(nv, mls, ch)
        mls_h = [oracle.fold_left_lerp_inplace(m, len(m), 0, nv, ch) for m in mls_h]
        if nv > 1:
            eq_d = be.fold_partial_eq_ind(nv - 1, eq_d)
            eq_h = oracle.fold_partial_eq_ind(eq_h)
    assert be._tail is None and (used_tail or n_vars == 1)
    for d, h in zip(mls, mls_h):
        assert d.evals.len() == 1 and _same(hal.to_host(d.evals), h)


def test_layer_calls_are_refused_while_a_tail_runs(hal, oracle):
    import binius_b200
    from binius_b200.hal import B200Backend, EqIndEvaluator, FoldedMultilinear

    be = B200Backend(hal, sumcheck_tail=True)
    n_vars = 4
    comps = u32_add_compositions()
    mls = [FoldedMultilinear(hal.to_device(oracle.rand_b128(4200 + t, 1 << n_vars)), 0) for t in range(5)]
    eq_d = be.tensor_product_full_query([3, 5, 7])
    be.sumcheck_compute_round_evals(n_vars, mls, [EqIndEvaluator(c) for c in comps], eq_d, [])
    assert be._tail is not None
    with pytest.raises(binius_b200.InputValidation):
        hal.to_host(eq_d)  # would otherwise queue behind the persistent kernel and deadlock
    for nv in range(n_vars, 0, -1):  # drive the tail to its end
        if nv < n_vars:
            be.sumcheck_compute_round_evals(nv, mls, [EqIndEvaluator(c) for c in comps], eq_d, [])
        be.sumcheck_fold_multilinears(nv, mls, 0x1234 + nv)
        if nv > 1:
            eq_d = be.fold_partial_eq_ind(nv - 1, eq_d)
    assert be._tail is None
    hal.to_host(eq_d)
