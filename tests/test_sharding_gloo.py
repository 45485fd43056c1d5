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


# ---- eq-ind (zerocheck) multilinear rounds sharded by the low variables ----------------------------------
class _FakeDevSlice:
    def __init__(self, arr):
        self.arr = arr

    def len(self):
        return len(self.arr)


class _FakeHalLayer:
    def to_device(self, host):
        return _FakeDevSlice(np.array(host, dtype=np.uint64, copy=True))

    def to_host(self, d):
        return d.arr.copy()


class _FakeBackend:
    """ComputationBackend-shaped stand-in (tests only): the four calls ShardedEqIndSumcheck makes, on the CPU oracle."""

    def __init__(self, orc):
        self.orc, self._l = orc, _FakeHalLayer()

    def tensor_product_full_query(self, q):
        return _FakeDevSlice(self.orc.tensor_expand(self.orc.to_arr([1] + [0] * ((1 << len(q)) - 1)), 0, list(q)))

    def sumcheck_compute_round_evals(self, n_vars, mls, evaluators, eq, finite_points=()):
        lo = min(ev.eval_point_indices().start for ev in evaluators)
        hi = max(ev.eval_point_indices().stop for ev in evaluators)
        codes = list(range(lo, hi))
        pts = [0 if c < 3 else finite_points[c - 3] for c in codes]
        vals = self.orc.sumcheck_round_evals(1, [m.evals.arr for m in mls], [len(m.evals.arr) for m in mls], [m.suffix_eval for m in mls], n_vars,
                                             eq.arr, [list(ev.composition.steps) for ev in evaluators],
                                             [list(ev.composition.leading_term().steps) for ev in evaluators], codes, pts)
        return [[row[k - lo] for k in ev.eval_point_indices()] for row, ev in zip(vals, evaluators)]

    def sumcheck_fold_multilinears(self, n_vars, mls, challenge):
        for m in mls:
            m.evals = _FakeDevSlice(self.orc.fold_left_lerp_inplace(m.evals.arr, len(m.evals.arr), m.suffix_eval, n_vars, challenge))
        return False

    def fold_partial_eq_ind(self, n_vars, eq):
        return eq if n_vars == 0 else _FakeDevSlice(self.orc.fold_partial_eq_ind(eq.arr))


def _eq_ind_instance(orc, n_vars):
    from binius_b200 import ArithCircuit as A

    rng = random.Random(77 + n_vars)
    mls = [orc.rand_b128(1200 + t, 1 << n_vars) for t in range(4)]
    comps = [A.var(0) * A.var(1) + A.var(2), A.var(1) * A.var(2) * A.var(3) + A.var(0) * A.constant(rng.getrandbits(128)), A.var(3) + A.var(0)]
    eq_ch = [rng.getrandbits(128) for _ in range(n_vars - 1)]
    pts = [rng.getrandbits(128)]  # one finite point: the degree-3 composition is evaluated at 1, infinity and this point
    ch = [rng.getrandbits(128) for _ in range(n_vars)]
    return mls, comps, eq_ch, pts, ch


def _eq_ind_reference(orc, mls, n_vars, comps, eq_ch, pts, ch, first_known):
    """the unsharded rounds straight on the oracle"""
    from binius_b200.hal import EqIndEvaluator

    cur = [m.copy() for m in mls]
    eq = orc.tensor_expand(orc.to_arr([1] + [0] * ((1 << (n_vars - 1)) - 1)), 0, list(eq_ch))
    rounds = []
    for r in range(n_vars):
        nv = n_vars - r
        evs = [EqIndEvaluator(c, first_known and r == 0) for c in comps]
        lo = min(ev.eval_point_indices().start for ev in evs)
        hi = max(ev.eval_point_indices().stop for ev in evs)
        codes = list(range(lo, hi))
        vals = orc.sumcheck_round_evals(1, cur, [len(m) for m in cur], [0] * len(cur), nv, eq, [list(c.steps) for c in comps],
                                        [list(c.leading_term().steps) for c in comps], codes, [0 if c < 3 else pts[c - 3] for c in codes])
        rounds.append([[row[k - lo] for k in ev.eval_point_indices()] for row, ev in zip(vals, evs)])
        cur = [orc.fold_left_lerp_inplace(m, len(m), 0, nv, ch[r]) for m in cur]
        if nv > 1:
            eq = orc.fold_partial_eq_ind(eq) if len(eq) > 1 else eq
    return rounds, [orc.to_ints(m)[0] for m in cur]


def _eq_ind_worker(rank, world, port, n_vars, first_known, ret):
    import torch.distributed as dist

    from binius_b200 import sharding
    from oracle import binding as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mls, comps, eq_ch, pts, ch = _eq_ind_instance(orc, n_vars)
        exp_rounds, exp_final = _eq_ind_reference(orc, mls, n_vars, comps, eq_ch, pts, ch, first_known)
        sc = sharding.ShardedEqIndSumcheck(_FakeBackend(orc), mls, n_vars, comps, eq_ch, pts, world, rank, dist,
                                           have_first_round_eval_1s=first_known)
        for r in range(n_vars):
            assert sc.round_evals() == exp_rounds[r], f"round {r} rank {rank}"
            sc.fold(ch[r])
        assert sc.finish() == exp_final
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_vars,first_known", [(2, 6, False), (4, 5, True), (4, 2, False)])
def test_sharded_eq_ind_sumcheck_gloo(world, n_vars, first_known):
    import torch.multiprocessing as mp

    port = 33500 + random.randrange(2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_eq_ind_worker, args=(world, port, n_vars, first_known, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_sharded_eq_ind_single_process_is_identity(oracle):
    from binius_b200 import sharding

    n_vars = 4
    mls, comps, eq_ch, pts, ch = _eq_ind_instance(oracle, n_vars)
    exp_rounds, exp_final = _eq_ind_reference(oracle, mls, n_vars, comps, eq_ch, pts, ch, False)
    sc = sharding.ShardedEqIndSumcheck(_FakeBackend(oracle), mls, n_vars, comps, eq_ch, pts)
    for r in range(n_vars):
        assert sc.round_evals() == exp_rounds[r]
        sc.fold(ch[r])
    assert sc.finish() == exp_final
