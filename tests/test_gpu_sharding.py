"""GPU parity test of the sharded eq-ind (zerocheck) rounds with the real CUDA layer: the W rank-local instances of
binius_b200.sharding.ShardedEqIndSumcheck run one after the other on one device (the collective -- an XOR of the
ranks' scaled partial values -- is done by hand), round by round against the unsharded oracle.  The NCCL combine and
the replicated tail are covered by tests/test_sharding_gloo.py (gloo, oracle-backed backend) and by
`bench.py --gpus N` (`sharded_sumcheck.eq_ind`, real layer over NCCL)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hal():
    import binius_b200

    layer = binius_b200.B200Layer(0)
    yield layer
    layer.close()


@pytest.mark.parametrize("world,n_vars,first_known", [(4, 12, False), (8, 10, True), (2, 15, False)])
def test_sharded_eq_ind_rounds_on_the_device(hal, oracle, world, n_vars, first_known):
    from binius_b200 import sharding
    from binius_b200.hal import B200Backend
    from test_sharding_gloo import _eq_ind_instance, _eq_ind_reference

    mls, comps, eq_ch, pts, ch = _eq_ind_instance(oracle, n_vars)
    exp_rounds, _ = _eq_ind_reference(oracle, mls, n_vars, comps, eq_ch, pts, ch, first_known)
    be = B200Backend(hal)
    ranks = [sharding.ShardedEqIndSumcheck(be, mls, n_vars, comps, eq_ch, pts, world, g, None, have_first_round_eval_1s=first_known)
             for g in range(world)]
    lw = world.bit_length() - 1
    cur = [m.copy() for m in mls]
    for r in range(n_vars - lw):
        parts = [sc.round_evals_local() for sc in ranks]
        got = [[0] * len(row) for row in parts[0]]
        for p in parts:
            for c, row in enumerate(p):
                for k, v in enumerate(row):
                    got[c][k] ^= v
        assert got == exp_rounds[r], f"round {r}"
        cur = [oracle.fold_left_lerp_inplace(m, len(m), 0, n_vars - r, ch[r]) for m in cur]
        for sc in ranks:
            sc.fold_local(ch[r])
    # one survivor per rank and multilinear: element g of the oracle's folded vector
    for g, sc in enumerate(ranks):
        for t, ml in enumerate(sc.mls):
            assert np.array_equal(hal.to_host(ml.evals)[:1].reshape(-1), cur[t][g].reshape(-1)), (g, t)
