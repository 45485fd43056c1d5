"""GPU parity tests of the ring-switch partial evaluations (core/src/ring_switch/prove.rs:147-208) and
evaluate_partial_high (math/src/multilinear_extension.rs:253-293) against the oracle's fold_left."""
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


@pytest.mark.parametrize("lvl,n_vars,log_q", [(0, 20, 13), (0, 22, 15), (0, 14, 7), (0, 10, 3), (3, 16, 12), (5, 14, 12), (7, 12, 12)])
def test_evaluate_partial_high(hal, oracle, lvl, n_vars, log_q):
    """(0, 20, 13) / (0, 22, 15): B1 witness, 2^13 / 2^15 query elements, 128 outputs -> the tensor-core outer product;
    the others take the generic fold kernels"""
    from binius_b200.hal import B200Backend
    from binius_b200.ring_switch import CommittedWitness, evaluate_partial_high

    be = B200Backend(hal)
    rng = random.Random(n_vars * 10 + lvl)
    n_words = (1 << n_vars) >> (7 - lvl)
    packed = oracle.rand_b128(800 + n_vars + lvl, max(n_words, 1))
    q = [rng.getrandbits(128) for _ in range(log_q)]
    qe = be.tensor_product_full_query(q)
    qe_h = hal.to_host(qe)
    n_out = (1 << n_vars) >> log_q
    got = evaluate_partial_high(hal, CommittedWitness(hal.to_device(packed), lvl, n_vars), qe)
    assert np.array_equal(hal.to_host(got), oracle.fold_left(packed, lvl, qe_h, n_out))


def test_compute_partial_evals(hal, oracle):
    from binius_b200.hal import B200Backend
    from binius_b200.ring_switch import CommittedWitness, compute_partial_evals

    be = B200Backend(hal)
    rng = random.Random(4)
    # two B1 witnesses (kappa 7), one B8 (kappa 4), one B32 shorter than 2^kappa after the partial evaluation (cycled)
    specs = [(0, 20), (0, 21), (3, 17), (5, 13)]
    hosts = [oracle.rand_b128(850 + i, (1 << nv) >> (7 - lvl)) for i, (lvl, nv) in enumerate(specs)]
    wit = [CommittedWitness(hal.to_device(h), lvl, nv) for h, (lvl, nv) in zip(hosts, specs)]
    sfx13 = tuple(rng.getrandbits(128) for _ in range(13))
    sfx14 = tuple(rng.getrandbits(128) for _ in range(14))
    claims = [(0, sfx13, 7), (1, sfx14, 7), (2, sfx13, 4), (0, sfx13, 7), (3, sfx13, 2)]
    got = compute_partial_evals(be, wit, claims)
    for (ci, sfx, kappa), g in zip(claims, got):
        lvl, nv = specs[ci]
        qe = oracle.tensor_expand(oracle.to_arr([1] + [0] * ((1 << len(sfx)) - 1)), 0, list(sfx))
        pe = oracle.to_ints(oracle.fold_left(hosts[ci], lvl, qe, (1 << nv) >> len(sfx)))[: 1 << kappa]
        while len(pe) < (1 << kappa):
            pe = (pe * 2)[: 1 << kappa]
        assert g == pe
