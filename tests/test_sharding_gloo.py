"""CPU tests (gloo, world_size 2 and 4) of the multi-GPU orchestration in binius_b200/sharding.py.
The device layer is replaced by an oracle-backed stand-in with the same executor surface, so what is
tested is the host-side logic: low-variable sharding, XOR-combine of partial round evaluations, the
survivor gather and the replicated tail rounds -- against the unsharded oracle, round by round."""
import os
import random

import numpy as np
import pytest


class _FakeSlice:
    def __init__(self, arr):
        self.arr = arr

    def len(self):
        return len(self.arr)

    def split_half_mut(self):
        h = len(self.arr) // 2
        return _FakeSlice(self.arr[:h]), _FakeSlice(self.arr[h:])


class _FakeExec:
    def __init__(self, orc):
        self.orc = orc

    def bivariate_round_evals(self, mls, n_vars, pairs, coeff):
        return self.orc.bivariate_round_evals([m.arr for m in mls], n_vars, pairs, coeff)

    def extrapolate_line(self, lo, hi, z):
        lo.arr[:] = self.orc.extrapolate_line(lo.arr, hi.arr, z)


class _FakeLayer:
    """ComputeLayer-shaped stand-in (tests only) computing with the CPU oracle."""

    def __init__(self, orc):
        self.orc = orc

    def to_device(self, host):
        return _FakeSlice(np.array(host, dtype=np.uint64, copy=True))

    def to_host(self, d):
        return d.arr.copy()

    def execute(self, f):
        return f(_FakeExec(self.orc))


def _reference_transcript(orc, mls, n_vars, pairs, alphas, challenges):
    out = []
    cur = [m.copy() for m in mls]
    for r in range(n_vars):
        nv = n_vars - r
        out.append(tuple(orc.bivariate_round_evals(cur, nv, pairs, alphas[r])))
        cur = [orc.extrapolate_line(m[: len(m) // 2], m[len(m) // 2:], challenges[r]) for m in cur]
    return out, [orc.to_ints(m)[0] for m in cur]


def _worker(rank, world, port, n_vars, ret):
    import torch.distributed as dist

    from binius_b200 import sharding
    from oracle import binding as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = random.Random(5)
        m = 4
        mls = [orc.rand_b128(900 + t, 1 << n_vars) for t in range(m)]
        pairs = [(0, 1), (2, 3), (1, 1)]
        alphas = [rng.getrandbits(128) for _ in range(n_vars)]
        challenges = [rng.getrandbits(128) for _ in range(n_vars)]
        exp_rounds, exp_final = _reference_transcript(orc, mls, n_vars, pairs, alphas, challenges)
        sc = sharding.ShardedBivariateSumcheck(_FakeLayer(orc), mls, n_vars, pairs, world, rank, dist)
        for r in range(n_vars):
            assert sc.round_evals(alphas[r]) == exp_rounds[r], f"round {r} rank {rank}"
            sc.fold(challenges[r])
        assert sc.finish() == exp_final
        # shard / unshard are inverse; unit partition covers everything exactly once
        assert np.array_equal(sharding.unshard_low_vars([sharding.shard_low_vars(mls[0], world, g) for g in range(world)]), mls[0])
        cover = sorted(i for g in range(world) for i in sharding.shard_units(13, world, g))
        assert cover == list(range(13))
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_vars", [(2, 6), (4, 5)])
def test_sharded_sumcheck_gloo(world, n_vars):
    import torch.multiprocessing as mp

    port = 29500 + random.randrange(2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_vars, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_hostfield_matches_oracle(oracle):
    from binius_b200 import hostfield

    rng = random.Random(0)
    for _ in range(50):
        a, b = rng.getrandbits(128), rng.getrandbits(128)
        assert hostfield.mul(a, b) == oracle.mul(a, b)
    for k in range(1, 7):
        a, b = rng.getrandbits(1 << k), rng.getrandbits(1 << k)
        assert hostfield.mul(a, b, k) == oracle.mul(a, b, k)
    pt = [rng.getrandbits(128) for _ in range(3)]
    eq = oracle.to_ints(oracle.tensor_expand(oracle.to_arr([1] + [0] * 7), 0, pt))
    assert [hostfield.eq_ind_scalar(i, pt) for i in range(8)] == eq


def test_single_process_sharding_is_identity(oracle):
    from binius_b200 import sharding

    n_vars = 5
    mls = [oracle.rand_b128(950 + t, 1 << n_vars) for t in range(2)]
    sc = sharding.ShardedBivariateSumcheck(_FakeLayer(oracle), mls, n_vars, [(0, 1)])
    exp, fin = _reference_transcript(oracle, mls, n_vars, [(0, 1)], [3] * n_vars, [7] * n_vars)
    for r in range(n_vars):
        assert sc.round_evals(3) == exp[r]
        sc.fold(7)
    assert sc.finish() == fin


# ---- zerocheck univariate-skip round sharded over sub-cubes ---------------------------------------------
def _uni_instance(orc, n_vars, skip):
    from binius_b200 import ArithCircuit as A
    from test_oracle_univariate import pack_scalars

    rng = random.Random(31 + skip)
    levels = [0, 3, 0]
    cols = [orc.to_arr(pack_scalars([rng.getrandbits(1 << l) for _ in range(1 << n_vars)], l)) for l in levels]
    comps = [A.var(0) * A.var(1) + A.var(2), A.var(0) * A.var(1) * A.var(2), A.var(0) + A.var(2)]
    ch = [rng.getrandbits(128) for _ in range(n_vars - skip)]
    return levels, cols, comps, ch


def _oracle_evaluator(orc):
    from binius_b200.hal import _degree

    def evaluate(local, levels, n_vars_local, skip, comps, challenges, max_domain):
        eq = orc.tensor_expand(orc.to_arr([1] + [0] * ((1 << len(challenges)) - 1)), 0, challenges)
        return orc.zerocheck_univariate_evals_reference(local, levels, n_vars_local, skip, eq, [list(c.steps) for c in comps],
                                                        [_degree(c) for c in comps], max_domain)

    return evaluate


def _uni_worker(rank, world, port, n_vars, skip, ret):
    import torch.distributed as dist

    from binius_b200 import sharding
    from oracle import binding as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        levels, cols, comps, ch = _uni_instance(orc, n_vars, skip)
        max_domain = (3 << skip) + 2
        exp = _oracle_evaluator(orc)(cols, levels, n_vars, skip, comps, ch, max_domain)
        got = sharding.sharded_zerocheck_univariate_evals(_oracle_evaluator(orc), cols, levels, n_vars, skip, comps, ch, max_domain,
                                                          world, rank, dist)
        assert got == exp, f"rank {rank}"
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_vars,skip", [(2, 7, 3), (4, 6, 1), (2, 6, 4)])
def test_sharded_univariate_round_gloo(world, n_vars, skip):
    import torch.multiprocessing as mp

    port = 31500 + random.randrange(2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_uni_worker, args=(world, port, n_vars, skip, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_shard_subcubes_partitions_every_subcube(oracle):
    from binius_b200 import sharding
    from test_oracle_univariate import pack_scalars

    rng = random.Random(2)
    for lvl, skip, n_vars in ((0, 1, 9), (0, 6, 9), (3, 2, 6), (5, 0, 5)):
        vals = [rng.getrandbits(1 << lvl) for _ in range(1 << n_vars)]
        col = oracle.to_arr(pack_scalars(vals, lvl))
        K = 1 << skip
        for world in (2, 4):
            for g in range(world):
                mine = [v for s in range(g, (1 << n_vars) // K, world) for v in vals[s * K:(s + 1) * K]]
                exp = oracle.to_arr(pack_scalars(mine, lvl))
                assert np.array_equal(sharding.shard_subcubes(col, lvl, skip, world, g)[: len(exp)], exp)
