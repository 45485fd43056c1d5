"""Oracle of the zerocheck univariate-skip round (oracle/univariate.c) checked against independent
restatements: Lagrange-basis identities, the skip = 0 degenerate case (plain eq-weighted sums), linearity
(a degree-1 "composition" = an inner product with L(x_i) (x) eq), and the reference's own test setting
(core/src/protocols/sumcheck/prove/univariate.rs:806-915: zero-product B1 multilinears), on which the
definition and the reference's evaluate-then-extrapolate route (univariate.rs:565-640) must coincide."""
import random

import numpy as np
import pytest


def pack_scalars(vals, lvl):
    """2^(7-lvl) scalars of 2^lvl bits per B128 word, low limb first (binary_field.rs:600-607)"""
    per = 1 << (7 - lvl)
    bits = 1 << lvl
    words = []
    for w in range((len(vals) + per - 1) // per):
        x = 0
        for j, v in enumerate(vals[w * per:(w + 1) * per]):
            x |= v << (j * bits)
        words.append(x)
    return words


def zero_product_columns(rng, n_vars, degree):
    """generate_zero_product_multilinears (core test_utils): B1 columns whose product vanishes everywhere"""
    n = 1 << n_vars
    cols = [[rng.getrandbits(1) for _ in range(n)] for _ in range(degree)]
    for i in range(n):
        cols[rng.randrange(degree)][i] = 0
    return cols


def product(ix):
    steps = [("var", i) for i in ix]
    acc = 0
    for t in range(1, len(ix)):
        steps.append(("mul", acc, t))
        acc = len(steps) - 1
    return steps


def test_lagrange_basis_identities(oracle):
    for k in (0, 1, 3, 5):
        n = 1 << k
        for u in range(n):  # L_t(u) = [t == u]
            assert oracle.lagrange_evals(k, u) == [int(t == u) for t in range(n)]
        for x in (n, n + 1, 0xAB, 0xFF):  # partition of unity, and values stay in B8
            ev = oracle.lagrange_evals(k, x)
            acc = 0
            for v in ev:
                acc ^= v
                assert v < 256
            assert acc == 1
        # extrapolating f(t) = t (degree 1) reproduces f
        if k >= 1:
            ev = oracle.lagrange_evals(k, 0xC5)
            acc = 0
            for t, v in enumerate(ev):
                acc ^= oracle.mul(v, t)
            assert acc == 0xC5


def test_skip_zero_is_a_plain_weighted_sum(oracle):
    rng = random.Random(5)
    n_vars = 6
    cols = [[rng.getrandbits(8) for _ in range(1 << n_vars)] for _ in range(3)]
    mls = [oracle.to_arr(pack_scalars(c, 3)) for c in cols]
    eq = [rng.getrandbits(128) for _ in range(1 << n_vars)]
    comp = product([0, 1, 2])
    got = oracle.zerocheck_univariate_evals(mls, [3, 3, 3], n_vars, 0, oracle.to_arr(eq), [comp], 4)
    exp = 0
    for s in range(1 << n_vars):
        exp ^= oracle.mul(eq[s], oracle.mul(oracle.mul(cols[0][s], cols[1][s]), cols[2][s]))
    assert got == [[exp] * 3]


@pytest.mark.parametrize("lvl", [0, 3, 5])
def test_linear_composition_is_an_inner_product(oracle, lvl):
    rng = random.Random(lvl)
    n_vars, skip = 8, 3
    K = 1 << skip
    col = [rng.getrandbits(1 << lvl) for _ in range(1 << n_vars)]
    ml = oracle.to_arr(pack_scalars(col, lvl))
    eq = [rng.getrandbits(128) for _ in range(1 << (n_vars - skip))]
    got = oracle.zerocheck_univariate_evals([ml], [lvl], n_vars, skip, oracle.to_arr(eq), [[("var", 0)]], 3 * K)[0]
    for i in (0, 1, K, 2 * K - 1):
        lag = oracle.lagrange_evals(skip, K + i)
        weights = [oracle.mul(eq[s], lag[t]) for s in range(len(eq)) for t in range(K)]
        assert got[i] == oracle.inner_product(ml, lvl, oracle.to_arr(weights))


@pytest.mark.parametrize("skip", [0, 1, 2, 4])
def test_reference_setting_definition_equals_extrapolation(oracle, skip):
    """On a true zerocheck instance the reference's route (evaluate at (deg-1)*2^k points, prepend zeros,
    interpolate, extend) equals the definition at every point of the domain."""
    rng = random.Random(100 + skip)
    n_vars = 6
    cols = zero_product_columns(rng, n_vars, 2) + zero_product_columns(rng, n_vars, 3) + zero_product_columns(rng, n_vars, 4)
    mls = [oracle.to_arr(pack_scalars(c, 0)) for c in cols]
    comps = [product([0, 1]), product([2, 3, 4]), product([5, 6, 7, 8])]
    degrees = [2, 3, 4]
    ch = [rng.getrandbits(128) for _ in range(n_vars - skip)]
    eq = oracle.tensor_expand(oracle.to_arr([1] + [0] * ((1 << len(ch)) - 1)), 0, ch)
    max_domain = 5 << skip
    full = oracle.zerocheck_univariate_evals(mls, [0] * 9, n_vars, skip, eq, comps, max_domain)
    ref = oracle.zerocheck_univariate_evals_reference(mls, [0] * 9, n_vars, skip, eq, comps, degrees, max_domain)
    assert all(len(v) == 4 << skip for v in full)
    assert full == ref
    assert skip == 0 or any(v != 0 for v in full[0])


def test_extrapolation_differs_off_a_true_instance(oracle):
    """With a composition that does not vanish on the cube the two routes differ beyond deg*2^k points --
    the product follows the reference's route (zeros assumed on the skipped domain)."""
    rng = random.Random(9)
    n_vars, skip = 5, 2
    cols = [[rng.getrandbits(1) for _ in range(1 << n_vars)] for _ in range(2)]
    mls = [oracle.to_arr(pack_scalars(c, 0)) for c in cols]
    eq = oracle.rand_b128(3, 1 << (n_vars - skip))
    comps = [product([0, 1])]
    full = oracle.zerocheck_univariate_evals(mls, [0, 0], n_vars, skip, eq, comps, 16)[0]
    ref = oracle.zerocheck_univariate_evals_reference(mls, [0, 0], n_vars, skip, eq, comps, [2], 16)[0]
    assert full[:4] == ref[:4] and full[4:] != ref[4:]


@pytest.mark.parametrize("skip,log_domain", [(1, 3), (2, 4), (3, 5), (5, 7), (6, 8), (7, 8)])
def test_lagrange_form_equals_the_reference_ntt_route(oracle, skip, log_domain):
    """ntt_extrapolate (univariate.rs:642-678): the reference turns a sub-cube's values into novel-basis
    coefficients with an inverse additive NTT over B8 (coset 0, coset_bits = log_domain - skip) and evaluates
    them on the following cosets with forward NTTs; coset c covers the domain points c*2^skip + t.  The oracle
    (and the kernels) use the Lagrange form instead -- same polynomial, so the same values.  The NTT used here is
    the oracle's restatement of ntt/src/tests/reference.rs, pinned on its own in test_oracle_ops.py."""
    K = 1 << skip
    ntt = oracle.NTT(3, log_domain)
    rng = random.Random(1000 + skip)
    evals = np.array([rng.getrandbits(8) for _ in range(K)], np.uint8)
    coset_bits = log_domain - skip
    coeffs = ntt.inverse(evals, 3, 0, skip, 0, 0, coset_bits, 0)
    assert np.array_equal(ntt.forward(coeffs, 3, 0, skip, 0, 0, coset_bits, 0), evals)
    for coset in range(1, 1 << coset_bits):
        ext = ntt.forward(coeffs, 3, 0, skip, 0, coset, coset_bits, 0)
        for t in sorted({0, 1, K // 2, K - 1}):
            lag = oracle.lagrange_evals(skip, coset * K + t)
            acc = 0
            for u in range(K):
                acc ^= oracle.mul(lag[u], int(evals[u]))
            assert acc == int(ext[t]), (coset, t)


@pytest.mark.parametrize("skip,degree,log_domain", [(3, 2, 6), (2, 4, 6), (4, 2, 7), (5, 4, 8), (0, 2, 3)])
def test_round_eval_extension_equals_the_reference_ntt_route(oracle, skip, degree, log_domain):
    """extrapolate_round_evals (univariate.rs:565-640) for a power-of-two number of evaluations, where
    OddInterpolate reduces to an inverse NTT: prepend 2^skip zeros, inverse NTT over the first deg*2^skip domain
    points, zero-pad the novel-basis coefficients to the domain, forward NTT, drop the skipped prefix."""
    K = 1 << skip
    n = degree * K
    rng = random.Random(77 + skip)
    stag = [rng.getrandbits(128) for _ in range(n - K)]
    ntt = oracle.NTT(3, log_domain)
    j = n.bit_length() - 1
    coeffs = ntt.inverse(oracle.to_arr([0] * K + stag), 7, 0, j, 0, 0, log_domain - j, 0)
    padded = np.zeros((1 << log_domain, 2), np.uint64)
    padded[:n] = coeffs
    full = oracle.to_ints(ntt.forward(padded, 7, 0, log_domain, 0, 0, 0, 0))
    assert full[:K] == [0] * K and full[K:n] == stag
    for max_domain in (1 << log_domain, n + 3):
        assert oracle.extrapolate_round_evals(stag, skip, degree, max_domain) == full[K:max_domain]


@pytest.mark.parametrize("skip,n_vars,threads", [(3, 8, 1), (4, 9, 3), (6, 10, 2), (7, 10, 3)])
def test_threaded_cpu_arm_matches_the_oracle(oracle, skip, n_vars, threads):
    """oracle/cpu_univariate.c (the timed CPU baseline of the univariate-skip round) against the definition."""
    rng = random.Random(500 + skip)
    m = 6
    cols = [oracle.to_arr(pack_scalars([rng.getrandbits(1) for _ in range(1 << n_vars)], 0)) for _ in range(m)]
    comps = [[("var", 0), ("var", 1), ("mul", 0, 1), ("var", 2), ("add", 2, 3), ("var", 5), ("add", 4, 5)],
             [("var", 3), ("const", 0x1D), ("mul", 0, 1), ("var", 4), ("var", 4), ("mul", 3, 4), ("add", 2, 5), ("const", 7), ("add", 6, 7)]]
    eq = oracle.rand_b128(9, 1 << (n_vars - skip))
    K = 1 << skip
    got, _ = oracle.cpu_univariate_b1(cols, n_vars, skip, eq, comps, K, threads)
    exp = oracle.zerocheck_univariate_evals(cols, [0] * m, n_vars, skip, eq, comps, 2 * K)
    assert got == exp
