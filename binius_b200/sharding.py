"""Multi-GPU sharding of the sumcheck hot path: one process per GPU, every round local.

HighToLow folding pairs index i with i + 2^(n-1), so partitioning every multilinear by its LOW
log2(W) index bits (rank g holds the compact array of elements with index = g mod W) keeps every
fold and every round-evaluation pass local to a GPU for the first n - log2(W) rounds.  Per round each
rank produces partial sums (<= 3 B128 per composition); GF(2^k) addition is XOR, which NCCL cannot
reduce, so the partials are all-gathered (48 B per rank) and XOR-ed locally -- they are needed on the
host for Fiat-Shamir anyway.  When a single element per rank is left, one all-gather brings the W
survivors together and the last log2(W) rounds run on those W elements (replicated, on the host).

The compute object is any ComputeLayer-shaped layer (`B200Layer` in production; the CPU tests drive the
same orchestration over gloo with an oracle-backed stand-in).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import hostfield


def shard_low_vars(host: np.ndarray, world: int, rank: int) -> np.ndarray:
    """elements with index = rank (mod world): the rank's sub-multilinear over the high variables"""
    assert world & (world - 1) == 0 and host.shape[0] % world == 0
    return np.ascontiguousarray(host[rank::world])


def unshard_low_vars(shards: Sequence[np.ndarray]) -> np.ndarray:
    world = len(shards)
    out = np.empty((shards[0].shape[0] * world,) + shards[0].shape[1:], dtype=shards[0].dtype)
    for g, s in enumerate(shards):
        out[g::world] = s
    return out


def xor_all_gather(values: Sequence[int], dist=None, device="cpu") -> List[int]:
    """XOR-combine a short vector of B128 scalars over all ranks (all_gather + local XOR)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(values)
    import torch

    t = torch.tensor([[v & 0x7FFFFFFFFFFFFFFF, (v >> 63) & 0x7FFFFFFFFFFFFFFF, v >> 126] for v in values], dtype=torch.int64, device=device)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    acc = [0] * len(values)
    for o in outs:
        for i, row in enumerate(o.tolist()):
            acc[i] ^= row[0] | (row[1] << 63) | (row[2] << 126)
    return acc


def gather_elements(local: np.ndarray, dist=None, device="cpu") -> np.ndarray:
    """all-gather the per-rank survivor elements ((k,2) uint64 each) in rank order -> (W, k, 2)"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local[None]
    import torch

    t = torch.from_numpy(local.view(np.int64).copy()).to(device)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return np.stack([o.cpu().numpy().view(np.uint64) for o in outs])


class ShardedBivariateSumcheck:
    """The prover-side data plane of v3::BivariateSumcheckProver
    (reference core/src/protocols/sumcheck/v3/bivariate_product.rs:57-252) sharded over W ranks:
    `round_evals` = calculate_round_evals (:303-408), `fold` = receive_challenge/fold (:168-232)."""

    def __init__(self, layer, host_multilinears: Sequence[np.ndarray], n_vars: int, pairs: Sequence[Tuple[int, int]],
                 world: int = 1, rank: int = 0, dist=None, comm_device="cpu"):
        self.layer, self.n_vars, self.pairs = layer, n_vars, list(pairs)
        self.world, self.rank, self.dist, self.comm_device = world, rank, dist, comm_device
        self.log_w = world.bit_length() - 1
        assert 1 << self.log_w == world and self.log_w <= n_vars
        self.local_vars = n_vars - self.log_w
        self.dev = [layer.to_device(shard_low_vars(m, world, rank)) for m in host_multilinears]
        self.tail = None  # (m, W) python ints once the local variables are exhausted

    def round_evals(self, batch_coeff: int) -> Tuple[int, int]:
        """(y_1, y_inf) of the current round, combined over all ranks"""
        if self.local_vars > 0:
            nv = self.local_vars
            got = self.layer.execute(lambda ex: list(ex.bivariate_round_evals(self.dev, nv, self.pairs, batch_coeff)))
            y1, yinf = xor_all_gather(got, self.dist, self.comm_device)
            return y1, yinf
        # replicated tail on the W gathered survivors (host scalars)
        half = len(self.tail[0]) // 2
        y1 = yinf = 0
        pw = 1
        for a, b in self.pairs:
            s1 = sinf = 0
            for i in range(half):
                s1 ^= hostfield.mul(self.tail[a][half + i], self.tail[b][half + i])
                sinf ^= hostfield.mul(self.tail[a][i] ^ self.tail[a][half + i], self.tail[b][i] ^ self.tail[b][half + i])
            y1 ^= hostfield.mul(s1, pw)
            yinf ^= hostfield.mul(sinf, pw)
            pw = hostfield.mul(pw, batch_coeff)
        return y1, yinf

    def fold(self, challenge: int):
        if self.local_vars > 0:
            def op(ex):
                for i, d in enumerate(self.dev):
                    lo, hi = d.split_half_mut()
                    ex.extrapolate_line(lo, hi, challenge)
                    self.dev[i] = lo
                return []

            self.layer.execute(op)
            self.local_vars -= 1
            if self.local_vars == 0:
                local = np.concatenate([self.layer.to_host(d) for d in self.dev])  # (m, 2)
                allv = gather_elements(local, self.dist, self.comm_device)  # (W, m, 2)
                self.tail = [[int(allv[g, t, 0]) | (int(allv[g, t, 1]) << 64) for g in range(self.world)] for t in range(len(self.dev))]
        else:
            new = []
            for vals in self.tail:
                half = len(vals) // 2
                new.append([vals[i] ^ hostfield.mul(vals[i] ^ vals[half + i], challenge) for i in range(half)])
            self.tail = new

    def finish(self) -> List[int]:
        """the fully folded value of every multilinear (bivariate_product.rs:245-252)"""
        assert self.tail is not None and all(len(v) == 1 for v in self.tail)
        return [v[0] for v in self.tail]


def _eval_circuit(circuit, q: Sequence[int]) -> int:
    """ArithCircuit (math/src/arith_expr.rs:367-383) on host scalars: only the replicated tail rounds of a
    sharded sumcheck use it (W / 2 hypercube points per round)."""
    tmp: List[int] = []
    for st in circuit.steps:
        k = st[0]
        if k == "add":
            v = tmp[st[1]] ^ tmp[st[2]]
        elif k == "mul":
            v = hostfield.mul(tmp[st[1]], tmp[st[2]])
        elif k == "pow":
            v = hostfield.pow_(tmp[st[1]], st[2])
        elif k == "const":
            v = st[1]
        else:
            v = q[st[1]]
        tmp.append(v)
    return tmp[-1] if tmp else 0


class ShardedEqIndSumcheck:
    """The data plane of the zerocheck / eq-ind multilinear rounds (reference EqIndSumcheckProver,
    core/src/protocols/sumcheck/prove/eq_ind.rs:300-470: `execute` -> backend.sumcheck_compute_round_evals with the
    evaluator of :646-731, `fold` -> sumcheck_fold_multilinears + fold_partial_eq_ind, prove/common.rs:13-73), HighToLow,
    sharded over W ranks by the LOW log2(W) variables like ShardedBivariateSumcheck.

    The eq-indicator is a tensor product, E[i] = eq(low bits of i; r_low) * E_high[high bits of i], and the round
    values are linear in E: rank g runs the ordinary rounds of its sub-instance (its elements i = g mod W, weighted by
    the expansion of the HIGH challenges only) and scales its partial values by the scalar eq(bits(g); r_low).  Per
    round one all-gather of n_compositions * n_points B128 values; no data-plane exchange.  When one element per rank is
    left the W survivors are gathered and the last log2(W) rounds run replicated on host scalars.

    `backend` is ComputationBackend-shaped (`B200Backend` in production); `eq_challenges` are the n_vars - 1 challenges
    whose expansion `Evaluator::eq_ind_partial_eval()` holds in round 0 (variables 0 .. n_vars-2; the round variable's
    own factor is applied by the prover outside the backend, eq_ind.rs:560-640).  Values returned per composition are at
    the evaluator's points 1 (unless `have_first_round_eval_1s` in round 0), infinity, then `finite_points`."""

    def __init__(self, backend, host_multilinears: Sequence[np.ndarray], n_vars: int, compositions, eq_challenges: Sequence[int],
                 finite_points: Sequence[int] = (), world: int = 1, rank: int = 0, dist=None, comm_device="cpu",
                 have_first_round_eval_1s: bool = False):
        from .hal import EqIndEvaluator, FoldedMultilinear

        assert len(eq_challenges) == max(n_vars - 1, 0)
        self.backend, self.n_vars, self.compositions = backend, n_vars, list(compositions)
        self.world, self.rank, self.dist, self.comm_device = world, rank, dist, comm_device
        self.log_w = world.bit_length() - 1
        assert 1 << self.log_w == world and self.log_w <= n_vars
        self.finite_points = list(finite_points)
        self.first_known = have_first_round_eval_1s
        self._Evaluator = EqIndEvaluator
        self.eq_challenges = list(eq_challenges)
        self.remaining = n_vars
        self.mls = [FoldedMultilinear(backend._l.to_device(shard_low_vars(m, world, rank))) for m in host_multilinears]
        self.tail = None
        lw = self.log_w
        if n_vars > lw:
            # variables lw .. n_vars-2 are local; an instance with a single local variable has the empty expansion [1]
            self.eq = backend.tensor_product_full_query(self.eq_challenges[lw:])
            self.scale = hostfield.eq_ind_scalar(rank, self.eq_challenges[:lw])
        else:
            self._gather()

    def _evaluators(self):
        first = self.first_known and self.remaining == self.n_vars
        return [self._Evaluator(c, first) for c in self.compositions]

    def _gather(self):
        layer = self.backend._l
        local = np.concatenate([layer.to_host(ml.evals)[:1] for ml in self.mls]) if self.mls else np.zeros((0, 2), np.uint64)
        allv = gather_elements(local, self.dist, self.comm_device)  # (W, m, 2)
        self.tail = [[int(allv[g, t, 0]) | (int(allv[g, t, 1]) << 64) for g in range(self.world)] for t in range(len(self.mls))]

    def _tail_round_evals(self) -> List[List[int]]:
        nv = self.remaining
        half = 1 << (nv - 1)
        eq = [hostfield.eq_ind_scalar(i, self.eq_challenges[:nv - 1]) for i in range(half)]
        out = []
        for ev in self._evaluators():
            lead = ev.composition.leading_term()
            row = []
            for code in ev.eval_point_indices():
                acc = 0
                for i in range(half):
                    lo = [v[i] for v in self.tail]
                    hi = [v[half + i] for v in self.tail]
                    if code == 1:
                        val = _eval_circuit(ev.composition, hi)
                    elif code == 2:
                        val = _eval_circuit(lead, [a ^ b for a, b in zip(lo, hi)])
                    else:
                        z = self.finite_points[code - 3]
                        val = _eval_circuit(ev.composition, [a ^ hostfield.mul(a ^ b, z) for a, b in zip(lo, hi)])
                    acc ^= hostfield.mul(val, eq[i])
                row.append(acc)
            out.append(row)
        return out

    def round_evals_local(self) -> List[List[int]]:
        """this rank's share of the round values (already scaled by eq(bits(rank); r_low)): their XOR over the ranks is
        the round's value"""
        assert self.tail is None and self.remaining > self.log_w
        part = self.backend.sumcheck_compute_round_evals(self.remaining - self.log_w, self.mls, self._evaluators(), self.eq,
                                                         self.finite_points)
        return [[hostfield.mul(self.scale, v) for v in row] for row in part]

    def round_evals(self) -> List[List[int]]:
        """RoundEvals of every composition, combined over all ranks"""
        if self.tail is not None:
            return self._tail_round_evals()
        part = self.round_evals_local()
        flat = [v for row in part for v in row]
        tot = xor_all_gather(flat, self.dist, self.comm_device) if flat else []
        out, k = [], 0
        for row in part:
            out.append(tot[k:k + len(row)])
            k += len(row)
        return out

    def fold_local(self, challenge: int):
        """fold of this rank's sub-instance (no communication)"""
        assert self.tail is None and self.remaining > self.log_w
        nv_local = self.remaining - self.log_w
        self.backend.sumcheck_fold_multilinears(nv_local, self.mls, challenge)
        if nv_local > 1:
            self.eq = self.backend.fold_partial_eq_ind(nv_local - 1, self.eq)
        self.remaining -= 1

    def fold(self, challenge: int):
        if self.tail is None:
            self.fold_local(challenge)
            if self.remaining == self.log_w:
                self._gather()
        else:
            new = []
            for vals in self.tail:
                half = len(vals) // 2
                new.append([vals[i] ^ hostfield.mul(vals[i] ^ vals[half + i], challenge) for i in range(half)])
            self.tail = new
            self.remaining -= 1

    def finish(self) -> List[int]:
        """the fully folded value of every multilinear (eq_ind.rs `finish`: multilinear_evals)"""
        assert self.tail is not None and self.remaining == 0
        return [v[0] for v in self.tail]


def shard_units(n_units: int, world: int, rank: int) -> range:
    """contiguous block partition of independent units (NTT batch columns, sumcheck instances)"""
    per, rem = divmod(n_units, world)
    start = rank * per + min(rank, rem)
    return range(start, start + per + (1 if rank < rem else 0))


# ---- zerocheck univariate-skip round (core/src/protocols/sumcheck/prove/univariate.rs:235-500) -------------
# The round evaluations are XOR-sums over sub-cubes s, and eq[s] = eq(low bits of s; r_low) * eq(high bits; r_high),
# so rank g takes the sub-cubes s = g (mod W) -- the same low-variable partition the multilinear rounds use
# afterwards --, evaluates them against the expansion of the HIGH challenges only, and scales its partial
# result by the scalar eq(bits(g), r_low).  The reference's domain extension (extrapolate_round_evals) is
# linear, so it commutes with the combine.  One all-gather of n_compositions * n_points B128 values.
def shard_subcubes(packed: np.ndarray, tower_level: int, skip_rounds: int, world: int, rank: int) -> np.ndarray:
    """sub-cubes s = rank (mod world) of a packed sub-field column ((n, 2) uint64 B128 words), compacted"""
    assert world & (world - 1) == 0
    cube_bits = (1 << skip_rounds) << tower_level
    raw = np.ascontiguousarray(packed).view(np.uint8).reshape(-1)
    if cube_bits % 8 == 0:
        mine = raw.reshape(-1, cube_bits // 8)[rank::world].reshape(-1)
    else:
        bits = np.unpackbits(raw, bitorder="little").reshape(-1, cube_bits)[rank::world].reshape(-1)
        mine = np.packbits(bits, bitorder="little")
    out = np.zeros(max((len(mine) + 15) // 16, 1) * 16, np.uint8)
    out[: len(mine)] = mine
    return out.view(np.uint64).reshape(-1, 2)


def sharded_zerocheck_univariate_evals(evaluate, columns: Sequence[np.ndarray], tower_levels: Sequence[int], n_vars: int,
                                       skip_rounds: int, compositions, zerocheck_challenges: Sequence[int], max_domain_size: int,
                                       world: int = 1, rank: int = 0, dist=None, device="cpu") -> List[List[int]]:
    """`evaluate(local_columns, tower_levels, n_vars_local, skip_rounds, compositions, challenges, max_domain_size)`
    returns the round evals of the rank-local instance (`device_univariate_evaluator` in production)."""
    lw = world.bit_length() - 1
    if n_vars - skip_rounds < lw:
        raise ValueError("fewer sub-cubes than ranks")
    low, high = list(zerocheck_challenges[:lw]), list(zerocheck_challenges[lw:])
    local = [shard_subcubes(c, l, skip_rounds, world, rank) for c, l in zip(columns, tower_levels)]
    part = evaluate(local, list(tower_levels), n_vars - lw, skip_rounds, compositions, high, max_domain_size)
    scale = hostfield.eq_ind_scalar(rank, low)
    flat = [hostfield.mul(scale, v) for row in part for v in row]
    tot = xor_all_gather(flat, dist, device) if flat else []
    n_out = len(part[0]) if part else 0
    return [tot[c * n_out:(c + 1) * n_out] for c in range(len(part))]


def device_univariate_evaluator(backend):
    """the rank-local evaluation on this rank's GPU (binius_b200.hal.zerocheck_univariate_evals)"""
    from .hal import TransparentMultilinear, zerocheck_univariate_evals

    def evaluate(local, levels, n_vars_local, skip, compositions, challenges, max_domain_size):
        layer = backend._l
        devs = [layer.to_device(c) for c in local]
        out = zerocheck_univariate_evals(backend, [TransparentMultilinear(d, l, n_vars_local) for d, l in zip(devs, levels)],
                                         compositions, challenges, skip, max_domain_size)
        layer.dev_free(out.partial_eq_ind_evals)
        for d in devs:
            layer.dev_free(d)
        return out.round_evals

    return evaluate
