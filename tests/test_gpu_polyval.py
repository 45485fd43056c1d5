"""GPU parity tests of the POLYVAL-facing GKR grand-product data plane: basis-change kernel,
GrandProductWitness layers and the eq-ind round values of the layer sumchecks, computed on the TOWER kernels and
compared bit for bit with the oracle's Montgomery arithmetic in BinaryField128bPolyval (oracle/polyval.c).
Sizes follow crates/core/benches/prodcheck.rs:41 (n_vars 12, 16, 20)."""
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


def test_linear_map_kernel(hal, oracle):
    from binius_b200 import polyval as pv

    t2p, p2t = pv.basis_change()
    for n in (1, 7, 1 << 12, (1 << 16) + 3):
        x = oracle.rand_b128(60 + n % 13, n)
        d = hal.to_device(x)
        out = hal.dev_alloc(n)
        pv.linear_map(hal, d, out, t2p)
        assert np.array_equal(hal.to_host(out), oracle.linear_map(t2p, x))
        pv.linear_map(hal, out, out, p2t)  # in place, back to the tower basis
        assert np.array_equal(hal.to_host(out), x)
    rnd = [random.Random(5).getrandbits(128) for _ in range(128)]  # an arbitrary (singular or not) linear map
    x = oracle.rand_b128(99, 1000)
    d, out = hal.to_device(x), hal.dev_alloc(1000)
    pv.linear_map(hal, d, out, rnd)
    assert np.array_equal(hal.to_host(out), oracle.linear_map(rnd, x))


@pytest.mark.parametrize("n_vars", [1, 5, 12, 16, 20])
def test_grand_product_witness_layers(hal, oracle, n_vars):
    from binius_b200 import polyval as pv

    t2p, _ = pv.basis_change()
    x_tower = oracle.rand_b128(300 + n_vars, 1 << n_vars)
    x_pv = oracle.linear_map(t2p, x_tower)  # what the prover hands over after convert_witnesses_to_fast_ext
    exp = oracle.polyval_gpa_layers(x_pv, n_vars)
    w = pv.GrandProductWitness(hal, n_vars, hal.to_device(x_pv))
    check = range(n_vars + 1) if n_vars <= 16 else (0, 1, 2, n_vars - 8, n_vars - 1, n_vars)
    for k in check:
        assert np.array_equal(w.layer_polyval(k), exp[k]), f"layer {k}"
    assert w.grand_product_evaluation() == oracle.to_ints(exp[-1])[0]


@pytest.mark.parametrize("n_vars", [3, 11, 13, 17])
def test_gpa_layer_sumcheck_rounds(hal, oracle, n_vars):
    """all rounds of the eq-ind sumcheck of one GPA layer (product of its two halves): round values in POLYVAL, folds"""
    from binius_b200 import polyval as pv
    from binius_b200.hal import B200Backend, FoldedMultilinear

    be = B200Backend(hal)
    rng = random.Random(n_vars)
    t2p, _ = pv.basis_change()
    layer_t = oracle.rand_b128(400 + n_vars, 2 << n_vars)  # tower basis; halves A | B
    a_p, b_p = oracle.linear_map(t2p, layer_t[: 1 << n_vars]), oracle.linear_map(t2p, layer_t[1 << n_vars:])
    eq_pt = [rng.getrandbits(128) for _ in range(n_vars - 1)]  # tower challenges
    eq_d = be.tensor_product_full_query(eq_pt)
    eq_p = oracle.linear_map(t2p, hal.to_host(eq_d))
    d_layer = hal.to_device(layer_t)
    mls = [FoldedMultilinear(d_layer.slice(0, 1 << n_vars), 0), FoldedMultilinear(d_layer.slice(1 << n_vars, 2 << n_vars), 0)]
    for rnd in range(min(n_vars, 5)):
        nv = n_vars - rnd
        cur = hal.dev_alloc(2 << nv)
        hal.copy_d2d(mls[0].evals, cur.slice(0, 1 << nv))
        hal.copy_d2d(mls[1].evals, cur.slice(1 << nv, 2 << nv))
        got = pv.gpa_round_evals(be, nv, cur, eq_d)
        assert got == oracle.polyval_gpa_round_evals(a_p, b_p, eq_p, nv), f"round {rnd}"
        ch = rng.getrandbits(128)
        be.sumcheck_fold_multilinears(nv, mls, ch)
        ch_p = pv.to_polyval(ch)
        half = 1 << (nv - 1)
        fold = lambda v: v[:half] ^ oracle.polyval_mul_vec(v[:half] ^ v[half:], oracle.to_arr([ch_p] * half))  # noqa: E731
        a_p, b_p = fold(a_p), fold(b_p)
        assert np.array_equal(oracle.linear_map(t2p, hal.to_host(mls[0].evals)), a_p)
        if nv > 1:
            eq_d = be.fold_partial_eq_ind(nv - 1, eq_d)
            eq_p = eq_p[: half // 2] ^ eq_p[half // 2: half] if half >= 2 else eq_p
