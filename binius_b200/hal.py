"""Host-side mirror of the old HAL `binius_hal::ComputationBackend` (reference
crates/hal/src/backend.rs:35-83) for device-resident multilinears, over the C ABI.

    tensor_product_full_query     backend.rs:42-45  -> math/src/tensor_prod_eq_ind.rs:94-101
    sumcheck_compute_round_evals  backend.rs:48-62  -> hal/src/sumcheck_round_calculation.rs:126-349
    sumcheck_fold_multilinears    backend.rs:65-75  -> hal/src/sumcheck_folding.rs:149-241
    evaluate_partial_high         backend.rs:78-82  -> math/src/multilinear_extension.rs:253-300

The reference's trait is generic over packed fields, multilinear trait objects and evaluator trait
objects; this backend covers what the prover instantiates it with: both `SumcheckMultilinear` variants
(hal/src/sumcheck_multilinear.rs:8-40; `Transparent` = packed sub-field multilinear folded into B128 at
its switchover round, `Folded` = B128 prefix + constant suffix), `ArithCircuit` compositions, both
evaluation orders, the eq-ind evaluator of core/src/protocols/sumcheck/prove/eq_ind.rs:646-731 and the
regular evaluator of prove/regular_sumcheck.rs:233-277.

Transparent multilinears before their switchover round: the reference evaluates
`subcube_partial_{high,low}_evals(tensor_query)` tile by tile on the fly to save host memory; here the
same partial evaluation (fold_left / fold_right of the packed sub-field matrix by the query expansion)
is written to device scratch for the round -- identical values, HBM is not the scarce resource.
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from .layer import (ArithCircuit, B200Layer, DevSlice, ExprEval, InputValidation, _u64_list, _u64x2)


@dataclass
class FoldedMultilinear:
    """SumcheckMultilinear::Folded { large_field_folded_evals, suffix_eval }: `evals` holds the stored
    prefix (its length may be shorter than 2^n_vars); the rest of the hypercube equals suffix_eval."""
    evals: DevSlice
    suffix_eval: int = 0


@dataclass
class TransparentMultilinear:
    """SumcheckMultilinear::Transparent { multilinear, switchover_round, const_suffix }: `evals` holds the
    2^n_vars sub-field scalars (tower level `tower_level`, 2^(7-level) per B128 word, low limb first --
    the MLEEmbeddingAdapter / SubfieldSlice layout, compute/src/memory.rs:257-281)."""
    evals: DevSlice
    tower_level: int
    n_vars: int
    switchover_round: int = 0
    const_suffix: Tuple[int, int] = (0, 0)  # (suffix_eval, suffix_len) hint


class EvaluationOrder(enum.IntEnum):
    """binius_math::EvaluationOrder"""
    LowToHigh = 0
    HighToLow = 1


def _degree(composition: ArithCircuit) -> int:
    cached = getattr(composition, "_degree", None)
    if cached is not None:
        return cached

    def deg(c, step):
        st = c.steps[step]
        if st[0] == "const":
            return 0
        if st[0] == "var":
            return 1
        if st[0] == "add":
            return max(deg(c, st[1]), deg(c, st[2]))
        if st[0] == "mul":
            return deg(c, st[1]) + deg(c, st[2])
        return deg(c, st[1]) * st[2]

    d = deg(composition, len(composition.steps) - 1)
    composition._degree = d
    return d


@dataclass
class RegularSumcheckEvaluator:
    """prove/regular_sumcheck.rs:221-277: plain sums of the composition, points 1..=degree
    (r(0) is derived from the claimed sum); no eq-indicator weighting."""
    composition: ArithCircuit

    def degree(self) -> int:
        return _degree(self.composition)

    def eval_point_indices(self) -> range:
        return range(1, self.degree() + 1)


@dataclass
class EqIndEvaluator:
    """The data of the eq-ind `Evaluator` (eq_ind.rs:646-662) that reaches the backend: the
    composition, and whether r(1) is already known (`have_first_round_eval_1s`)."""
    composition: ArithCircuit
    have_first_round_eval_1s: bool = False

    def degree(self) -> int:
        return _degree(self.composition)

    def eval_point_indices(self) -> range:
        # eq_ind.rs:667-671: skip r(1) in the first round when known; never evaluate at 0
        return range(2 if self.have_first_round_eval_1s else 1, self.degree() + 1)


@dataclass
class ZerocheckUnivariateEvalsOutput:
    """core/src/protocols/sumcheck/prove/univariate.rs:37-50"""
    round_evals: List[List[int]]
    skip_rounds: int
    remaining_rounds: int
    max_domain_size: int
    partial_eq_ind_evals: DevSlice


def domain_size(composition_degree: int, skip_rounds: int) -> int:
    """core/src/protocols/sumcheck/zerocheck.rs:131-133"""
    return composition_degree << skip_rounds


def zerocheck_univariate_evals(backend: "B200Backend", multilinears: Sequence[TransparentMultilinear],
                               compositions: Sequence[ArithCircuit], zerocheck_challenges: Sequence[int],
                               skip_rounds: int, max_domain_size: int) -> ZerocheckUnivariateEvalsOutput:
    """Mirror of `zerocheck_univariate_evals` (core/src/protocols/sumcheck/prove/univariate.rs:235-500) with
    FDomain = BinaryField8b: the round evaluations of every composition at the max_domain_size - 2^skip_rounds
    domain points after the skipped sub-cube, and the eq-indicator expansion of the remaining challenges.
    Multilinears are packed sub-field columns on the device (the MLEEmbeddingAdapter layout)."""
    if not multilinears:
        raise InputValidation("NumberOfVariablesMismatch: no multilinears")
    n_vars = multilinears[0].n_vars
    if any(ml.n_vars != n_vars for ml in multilinears):
        raise InputValidation("NumberOfVariablesMismatch")
    if skip_rounds > n_vars:
        raise InputValidation("TooManySkippedRounds")
    remaining_rounds = n_vars - skip_rounds
    if len(zerocheck_challenges) != remaining_rounds:
        raise InputValidation("IncorrectZerocheckChallengesLength")
    degrees = [_degree(c) for c in compositions]
    if max_domain_size < domain_size(max(degrees, default=0), skip_rounds):
        raise InputValidation("LagrangeDomainTooSmall")
    if max_domain_size > 256:
        raise InputValidation("DomainSizeTooLarge")
    L = backend._l
    partial_eq_ind_evals = backend.tensor_product_full_query(zerocheck_challenges)
    m, nc = len(multilinears), len(compositions)
    n_out = max_domain_size - (1 << skip_rounds)
    ptrs = (C.c_void_p * m)(*[ml.evals.ptr for ml in multilinears])
    lvls = (C.c_uint32 * m)(*[ml.tower_level for ml in multilinears])
    for ml in multilinears:
        if ml.evals.len() << (7 - ml.tower_level) < 1 << n_vars:
            raise InputValidation("multilinear buffer shorter than 2^n_vars scalars")
    comps = (C.c_void_p * max(nc, 1))(*[backend._compiled(c)[0].handle.value for c in compositions])
    degs = (C.c_uint32 * max(nc, 1))(*degrees)
    out = (C.c_uint64 * max(2 * nc * n_out, 2))()
    L._check(L._lib.b200_zerocheck_univariate_evals(L._ctx, ptrs, lvls, m, n_vars, skip_rounds, partial_eq_ind_evals.ptr,
                                                    partial_eq_ind_evals.len(), comps, degs, nc, max_domain_size, out))
    vals = [int(out[2 * i]) | (int(out[2 * i + 1]) << 64) for i in range(nc * n_out)]
    return ZerocheckUnivariateEvalsOutput([vals[c * n_out:(c + 1) * n_out] for c in range(nc)], skip_rounds,
                                          remaining_rounds, max_domain_size, partial_eq_ind_evals)


def zerocheck_univariate_evals_streamed(backend: "B200Backend", host_columns: Sequence[np.ndarray], multilinears: Sequence[TransparentMultilinear],
                                        compositions: Sequence[ArithCircuit], zerocheck_challenges: Sequence[int],
                                        skip_rounds: int, max_domain_size: int, log_chunks: int = 3) -> ZerocheckUnivariateEvalsOutput:
    """The same round with the witness still in HOST memory (pinned `host_columns`, one per multilinear): the columns are
    uploaded into `multilinears[j].evals` in 2^log_chunks row chunks on the side stream while the previous chunk is
    evaluated (b200_zerocheck_univariate_evals_streamed) -- the round values are XOR-sums over sub-cubes, so the chunks' values add up and the
    reference's domain extension is linear --, i.e. the witness upload hides behind the univariate round.  Afterwards
    the columns are resident on the device for the multilinear rounds."""
    if not multilinears or len(host_columns) != len(multilinears):
        raise InputValidation("one host column per multilinear")
    n_vars = multilinears[0].n_vars
    remaining = n_vars - skip_rounds
    if remaining < 0 or len(zerocheck_challenges) != remaining:
        raise InputValidation("IncorrectZerocheckChallengesLength")
    degrees = [_degree(c) for c in compositions]
    if max_domain_size < domain_size(max(degrees, default=0), skip_rounds):
        raise InputValidation("LagrangeDomainTooSmall")
    if max_domain_size > 256:
        raise InputValidation("DomainSizeTooLarge")
    for ml, h in zip(multilinears, host_columns):
        if ml.n_vars != n_vars or h.nbytes < max(((1 << n_vars) << ml.tower_level) // 8, 1):
            raise InputValidation("NumberOfVariablesMismatch")
    L = backend._l
    eq = backend.tensor_product_full_query(zerocheck_challenges)
    m, nc = len(multilinears), len(compositions)
    n_out = max_domain_size - (1 << skip_rounds)
    hosts = (C.c_void_p * m)(*[h.ctypes.data for h in host_columns])
    ptrs = (C.c_void_p * m)(*[ml.evals.ptr for ml in multilinears])
    lvls = (C.c_uint32 * m)(*[ml.tower_level for ml in multilinears])
    comps = (C.c_void_p * max(nc, 1))(*[backend._compiled(c)[0].handle.value for c in compositions])
    degs = (C.c_uint32 * max(nc, 1))(*degrees)
    out = (C.c_uint64 * max(2 * nc * n_out, 2))()
    L._check(L._lib.b200_zerocheck_univariate_evals_streamed(L._ctx, hosts, ptrs, lvls, m, n_vars, skip_rounds, eq.ptr, 1 << remaining, comps, degs, nc,
                                                             max_domain_size, max(0, log_chunks), out))
    vals = [int(out[2 * i]) | (int(out[2 * i + 1]) << 64) for i in range(nc * n_out)]
    return ZerocheckUnivariateEvalsOutput([vals[c * n_out:(c + 1) * n_out] for c in range(nc)], skip_rounds, remaining, max_domain_size, eq)


@dataclass
class PreparedUnivariateRound:
    """What `zerocheck_univariate_prepare` leaves behind for `zerocheck_univariate_finish`: the B8 values of every
    composition's non-linear part on the extrapolation domain (`store`, owned by this object until `release`)."""
    multilinears: Sequence[TransparentMultilinear]
    compositions: Sequence[ArithCircuit]
    skip_rounds: int
    max_domain_size: int
    store: Optional[DevSlice]
    prepared: bool
    owns_store: bool = True

    def release(self, backend: "B200Backend"):
        if self.store is not None and self.owns_store:
            backend._l.dev_free(self.store)
        self.store = None


def _uni_call_args(backend, multilinears, compositions):
    m, nc = len(multilinears), len(compositions)
    ptrs = (C.c_void_p * m)(*[ml.evals.ptr for ml in multilinears])
    lvls = (C.c_uint32 * m)(*[ml.tower_level for ml in multilinears])
    comps = (C.c_void_p * max(nc, 1))(*[backend._compiled(c)[0].handle.value for c in compositions])
    degs = (C.c_uint32 * max(nc, 1))(*[_degree(c) for c in compositions])
    return ptrs, lvls, comps, degs


def zerocheck_univariate_prepare(backend: "B200Backend", multilinears: Sequence[TransparentMultilinear], compositions: Sequence[ArithCircuit],
                                 skip_rounds: int, max_domain_size: int, host_columns: Optional[Sequence[np.ndarray]] = None,
                                 log_chunks: int = 3, arena_store: Optional[DevSlice] = None) -> PreparedUnivariateRound:
    """The challenge-independent half of `zerocheck_univariate_evals` (b200_zerocheck_univariate_prepare): sub-cube
    extrapolations and composition values on the extrapolation domain (univariate.rs:380-470 need the witness only), run
    BEFORE the zerocheck challenges exist -- in the reference's order (commit, then zerocheck) that is during the witness
    upload (`host_columns`: pinned host copies, uploaded chunk by chunk into `multilinears[j].evals` as in the streamed
    round) and the commitment.  Shapes outside the B8 fast path are left to `zerocheck_univariate_finish`'s fallback."""
    if not multilinears:
        raise InputValidation("NumberOfVariablesMismatch: no multilinears")
    n_vars = multilinears[0].n_vars
    if any(ml.n_vars != n_vars for ml in multilinears):
        raise InputValidation("NumberOfVariablesMismatch")
    if skip_rounds > n_vars:
        raise InputValidation("TooManySkippedRounds")
    degrees = [_degree(c) for c in compositions]
    if max_domain_size < domain_size(max(degrees, default=0), skip_rounds):
        raise InputValidation("LagrangeDomainTooSmall")
    if max_domain_size > 256:
        raise InputValidation("DomainSizeTooLarge")
    if host_columns is not None:
        if len(host_columns) != len(multilinears):
            raise InputValidation("one host column per multilinear")
        for ml, h in zip(multilinears, host_columns):
            if h.nbytes < max(((1 << n_vars) << ml.tower_level) // 8, 1):
                raise InputValidation("NumberOfVariablesMismatch")
    L = backend._l
    ptrs, lvls, comps, degs = _uni_call_args(backend, multilinears, compositions)
    m, nc = len(multilinears), len(compositions)
    n_store = int(L._lib.b200_zerocheck_univariate_store_elems(n_vars, skip_rounds, degs, nc))
    # the value store: a slice of the caller's device arena when given, else an allocation this object owns
    if arena_store is not None:
        if arena_store.len() < n_store:
            raise InputValidation("the store slice is smaller than b200_zerocheck_univariate_store_elems")
        store = arena_store.slice(0, n_store) if n_store else None
    else:
        store = L.dev_alloc(n_store) if n_store else None
    owns = arena_store is None
    hosts = (C.c_void_p * m)(*[h.ctypes.data for h in host_columns]) if host_columns is not None else None
    done = C.c_uint32()
    try:
        L._check(L._lib.b200_zerocheck_univariate_prepare(L._ctx, hosts, ptrs, lvls, m, n_vars, skip_rounds, comps, degs, nc, max_domain_size,
                                                          max(0, log_chunks), store.ptr if store else None, C.c_uint64(n_store), C.byref(done)))
    except Exception:
        if store is not None and owns:
            L.dev_free(store)
        raise
    prep = PreparedUnivariateRound(list(multilinears), list(compositions), skip_rounds, max_domain_size, store, bool(done.value), owns)
    if not prep.prepared:
        prep.release(backend)
    return prep


def zerocheck_univariate_finish(backend: "B200Backend", prep: PreparedUnivariateRound, zerocheck_challenges: Sequence[int]) -> ZerocheckUnivariateEvalsOutput:
    """The challenge-dependent half: weights the prepared values by the eq-indicator of `zerocheck_challenges`
    (b200_zerocheck_univariate_finish); values identical to `zerocheck_univariate_evals`, which it calls itself when the
    shape was not a prepared one."""
    if not prep.prepared:
        return zerocheck_univariate_evals(backend, prep.multilinears, prep.compositions, zerocheck_challenges, prep.skip_rounds, prep.max_domain_size)
    n_vars, skip = prep.multilinears[0].n_vars, prep.skip_rounds
    if len(zerocheck_challenges) != n_vars - skip:
        raise InputValidation("IncorrectZerocheckChallengesLength")
    L = backend._l
    eq = backend.tensor_product_full_query(zerocheck_challenges)
    ptrs, lvls, comps, degs = _uni_call_args(backend, prep.multilinears, prep.compositions)
    m, nc = len(prep.multilinears), len(prep.compositions)
    n_out = prep.max_domain_size - (1 << skip)
    out = (C.c_uint64 * max(2 * nc * n_out, 2))()
    L._check(L._lib.b200_zerocheck_univariate_finish(L._ctx, ptrs, lvls, m, n_vars, skip, eq.ptr, C.c_uint64(eq.len()), comps, degs, nc, prep.max_domain_size,
                                                     prep.store.ptr, C.c_uint64(prep.store.len()), out))
    vals = [int(out[2 * i]) | (int(out[2 * i + 1]) << 64) for i in range(nc * n_out)]
    return ZerocheckUnivariateEvalsOutput([vals[c * n_out:(c + 1) * n_out] for c in range(nc)], skip, n_vars - skip, prep.max_domain_size, eq)


class B200Backend:
    """ComputationBackend over one B200Layer."""

    def __init__(self, layer: B200Layer, sumcheck_tail: bool = False):
        self._l = layer
        self._exprs = {}
        self._unit = None
        self._ones2 = None
        # out-of-place results of the LowToHigh paths (fold_right_lerp, eq-ind halving) come from this pool and the
        # buffer they replace goes back to it: two buffers per multilinear ping-pong instead of one allocation per
        # round (buffers the caller handed in are never recycled: only pointers taken from the pool are given back)
        self._mine = {}
        self._free = []
        # persistent sumcheck tail (b200_sumcheck_tail_*): when the (composition, point, hypercube index) triples of a
        # round drop below `tail_threshold` (2^20: above it the tensor-core route of the separate calls is faster than
        # the per-lane products of the persistent kernel), the remaining rounds of an eq-ind prover -- all of them for
        # a cache-sized instance like BASELINE config #3 -- run in ONE kernel and the trait
        # calls of those rounds only read / post through its host-mapped mailboxes
        # (opt-in: while the tail kernel runs it owns the context's stream -- any other call on the layer is refused --,
        # so only a caller whose round loop does nothing else on the layer, like the prover's, should switch it on)
        self.tail_threshold = (1 << 20) if sumcheck_tail else 0
        self._tail = None

    def _take(self, n: int) -> DevSlice:
        n = max(n, 1)
        best = None
        for i, (_, cap) in enumerate(self._free):
            if cap >= n and (best is None or cap < self._free[best][1]):
                best = i
        if best is not None:
            ptr, cap = self._free.pop(best)
        else:
            ptr, cap = self._l.dev_alloc(n).ptr, n
        self._mine[ptr] = cap
        return DevSlice(ptr, n)

    def _give(self, s: DevSlice):
        cap = self._mine.pop(s.ptr, None)
        if cap is not None:  # one in-order stream: a later kernel that reuses the buffer runs after its last reader
            self._free.append((s.ptr, cap))

    def _compiled(self, circuit: ArithCircuit) -> ExprEval:
        hit = getattr(circuit, "_b200_compiled", None)
        if hit is not None and hit[0] is self:
            return hit[1]
        key = tuple(circuit.steps)
        if key not in self._exprs:
            self._exprs[key] = (self._l.compile_expr(circuit), self._l.compile_expr(circuit.leading_term()))
        circuit._b200_compiled = (self, self._exprs[key])
        return self._exprs[key]

    # -- backend.rs:42-45
    def tensor_product_full_query(self, query: Sequence[int]) -> DevSlice:
        out = self._l.dev_alloc(1 << len(query))
        self._l._check(self._l._lib.b200_tensor_product_full_query(self._l._ctx, _u64_list(query), len(query), out.ptr, out.len()))
        return out

    def _one(self) -> DevSlice:
        """the empty tensor query's expansion: a single 1 (MultilinearQuery::with_capacity(0))"""
        if self._unit is None:
            self._unit = self._l.dev_alloc(1)
            self._l.fill(self._unit, 1)
        return self._unit

    def _partial_eval(self, order: EvaluationOrder, ml: TransparentMultilinear, tensor_query: Optional[DevSlice]) -> DevSlice:
        """evaluate_partial_high / evaluate_partial_low of a packed sub-field multilinear by an expanded
        tensor query (math/src/multilinear_extension.rs:253-341 -> fold_left / fold_right, fold.rs:88-179)."""
        q = tensor_query if tensor_query is not None else self._one()
        n_scalars = 1 << ml.n_vars
        if q.len() == 0 or n_scalars % q.len():
            raise InputValidation("tensor query does not divide the multilinear")
        out = self._l.dev_alloc(n_scalars // q.len())
        fn = self._l._lib.b200_fold_left if order == EvaluationOrder.HighToLow else self._l._lib.b200_fold_right
        self._l._check(fn(self._l._ctx, ml.evals.ptr, ml.evals.len(), ml.tower_level, q.ptr, q.len(), out.ptr, out.len()))
        return out

    # -- backend.rs:48-62
    def sumcheck_compute_round_evals(self, n_vars: int, multilinears: Sequence[Union[FoldedMultilinear, TransparentMultilinear]],
                                     evaluators: Sequence[Union[EqIndEvaluator, RegularSumcheckEvaluator]],
                                     eq_ind_partial_evals: Optional[DevSlice] = None,
                                     finite_evaluation_points: Sequence[int] = (), *,
                                     evaluation_order: EvaluationOrder = EvaluationOrder.HighToLow,
                                     tensor_query: Optional[DevSlice] = None) -> List[List[int]]:
        """Returns RoundEvals per evaluator: the values at eval_point_indices() (1 = at 1, 2 = infinity,
        k >= 3 = finite_evaluation_points[k-3]); sumcheck_round_calculation.rs:156-163 length check.
        `eq_ind_partial_evals` is what Evaluator::eq_ind_partial_eval() returns (None for the regular
        evaluator); `tensor_query` is the expansion of the challenges received so far, needed while any
        multilinear is still Transparent."""
        if n_vars == 0:
            raise InputValidation("Computing round evaluations requires at least a single variable.")
        hi = max([ev.eval_point_indices().stop for ev in evaluators], default=0)
        if len(finite_evaluation_points) != max(hi - 3, 0):
            raise InputValidation("IncorrectNontrivialEvalPointsLength")
        weighted = any(isinstance(ev, EqIndEvaluator) for ev in evaluators)
        if weighted and not all(isinstance(ev, EqIndEvaluator) for ev in evaluators):
            raise InputValidation("evaluators of one call must be of one kind")
        if weighted and (eq_ind_partial_evals is None or eq_ind_partial_evals.len() != 1 << (n_vars - 1)):
            raise InputValidation("eq_ind_partial_evals must have 2^(n_vars-1) elements")
        lo = min([ev.eval_point_indices().start for ev in evaluators], default=1)
        codes = list(range(lo, hi))
        if not codes or not evaluators:
            return [[] for _ in evaluators]
        pts = [0 if c < 3 else finite_evaluation_points[c - 3] for c in codes]
        if self._tail is not None or self._tail_eligible(n_vars, multilinears, evaluators, weighted, evaluation_order, codes):
            return self._tail_round_evals(n_vars, multilinears, evaluators, eq_ind_partial_evals, codes, pts, lo)
        temps = []
        views = []
        for ml in multilinears:
            if isinstance(ml, TransparentMultilinear):
                d = self._partial_eval(evaluation_order, ml, tensor_query)
                if d.len() != 1 << n_vars:
                    raise InputValidation("transparent multilinear does not match n_vars and the tensor query")
                temps.append(d)
                views.append((d, ml.const_suffix[0] if ml.const_suffix[1] else 0))
            else:
                views.append((ml.evals, ml.suffix_eval))
        m = len(views)
        ptrs = (C.c_void_p * max(m, 1))(*[v[0].ptr for v in views])
        lens = (C.c_uint64 * max(m, 1))(*[min(v[0].len(), 1 << n_vars) for v in views])
        sfx = _u64_list([v[1] for v in views])
        compiled = [self._compiled(ev.composition) for ev in evaluators]
        comps = (C.c_void_p * len(evaluators))(*[c[0].handle.value for c in compiled])
        leads = (C.c_void_p * len(evaluators))(*[c[1].handle.value for c in compiled])
        ccodes = (C.c_uint32 * len(codes))(*codes)
        first = C.c_uint32()
        L = self._l
        L._check(L._lib.b200_results_reset(L._ctx))
        L._check(L._lib.b200_sumcheck_round_evals(L._ctx, int(evaluation_order), ptrs, lens, sfx, m, n_vars,
                                                  eq_ind_partial_evals.ptr if weighted else None, comps, leads,
                                                  len(evaluators), ccodes, _u64_list(pts), len(codes), C.byref(first)))
        total = len(evaluators) * len(codes)
        slots = (C.c_uint32 * total)(*range(first.value, first.value + total))
        out = (C.c_uint64 * (2 * total))()
        L._check(L._lib.b200_results_fetch(L._ctx, slots, total, out))  # synchronises: temps may go
        for d in temps:
            L.dev_free(d)
        vals = [int(out[2 * i]) | (int(out[2 * i + 1]) << 64) for i in range(total)]
        res = []
        for e, ev in enumerate(evaluators):
            rng = ev.eval_point_indices()
            res.append([vals[e * len(codes) + (k - lo)] for k in rng])
        return res

    # -- persistent tail -------------------------------------------------------------------------------------------
    def _tail_eligible(self, n_vars, multilinears, evaluators, weighted, order, codes) -> bool:
        if not self.tail_threshold or not weighted or order != EvaluationOrder.HighToLow or n_vars > 28:
            return False
        if any(not isinstance(ml, FoldedMultilinear) or ml.evals.len() != 1 << n_vars for ml in multilinears):
            return False
        n_vals = len(evaluators) * (codes[-1] + 1 - 1)  # the tail computes the points 1..hi-1 of every evaluator
        grid = n_vars - 1 > 5  # more than 32 hypercube points: the co-resident grid kernel
        if (grid and n_vals > 1024) or (not grid and (n_vals > 4096 or codes[0] != 1)):
            return False
        return n_vals << (n_vars - 1) <= self.tail_threshold

    def _tail_round_evals(self, n_vars, multilinears, evaluators, eq_ind, codes, pts, lo):
        L = self._l
        hi = codes[-1] + 1
        n_codes = hi - 1  # the tail's point list is 1..hi-1 in every round; a prover's first round starts at `lo` = 2
        if self._tail is None:
            compiled = [self._compiled(ev.composition) for ev in evaluators]
            m = len(multilinears)
            ptrs = (C.c_void_p * max(m, 1))(*[ml.evals.ptr for ml in multilinears])
            comps = (C.c_void_p * len(evaluators))(*[c[0].handle.value for c in compiled])
            leads = (C.c_void_p * len(evaluators))(*[c[1].handle.value for c in compiled])
            h = C.c_void_p()
            L._check(L._lib.b200_sumcheck_tail_start(L._ctx, ptrs, m, n_vars, eq_ind.ptr, comps, leads, len(evaluators),
                                                     (C.c_uint32 * n_codes)(*range(1, hi)), _u64_list([0] * (lo - 1) + list(pts)), n_codes, lo - 1,
                                                     C.byref(h)))
            self._tail = {"h": h, "n_vars": n_vars, "n_comp": len(evaluators), "hi": hi, "eq_ptr": eq_ind.ptr, "eq_pending": 0, "first": True}
        t = self._tail
        if t["n_vars"] != n_vars or t["n_comp"] != len(evaluators) or t["hi"] != hi or (lo != 1 and not t["first"]):
            raise InputValidation("the running sumcheck tail was started for a different round shape")
        t["first"] = False
        total = len(evaluators) * n_codes
        out = (C.c_uint64 * (2 * total))()
        L._check(L._lib.b200_sumcheck_tail_round_evals(t["h"], out))
        vals = [int(out[2 * i]) | (int(out[2 * i + 1]) << 64) for i in range(total)]
        return [[vals[e * n_codes + (k - 1)] for k in ev.eval_point_indices()] for e, ev in enumerate(evaluators)]

    def _tail_fold(self, n_vars, multilinears, challenge):
        L, t = self._l, self._tail
        if t["n_vars"] != n_vars:
            raise InputValidation("fold does not match the running sumcheck tail")
        L._check(L._lib.b200_sumcheck_tail_challenge(t["h"], _u64x2(challenge)))
        for ml in multilinears:
            ml.evals = ml.evals.slice(0, 1 << (n_vars - 1))
        t["n_vars"] -= 1
        t["eq_pending"] += 1  # the kernel halves the eq-indicator itself: the next fold_partial_eq_ind is a view change
        if t["n_vars"] == 0:
            self._tail = None
            L._check(L._lib.b200_sumcheck_tail_finish(t["h"]))
            self._tail_done_eq = t["eq_ptr"]
        return False

    # -- backend.rs:65-75: folds every multilinear by `challenge`; Folded ones by a single-variable lerp
    #    (in place for HighToLow, into a fresh buffer for LowToHigh), Transparent ones are materialised by
    #    `tensor_query` (which already includes `challenge`, prover_state.rs fold()) at their switchover
    #    round.  Entries of `multilinears` are replaced/truncated; returns any_transparent_left.
    def sumcheck_fold_multilinears(self, n_vars: int, multilinears: List[Union[FoldedMultilinear, TransparentMultilinear]],
                                   challenge: int, tensor_query: Optional[DevSlice] = None, *,
                                   evaluation_order: EvaluationOrder = EvaluationOrder.HighToLow) -> bool:
        L = self._l
        if self._tail is not None:
            return self._tail_fold(n_vars, multilinears, challenge)
        any_transparent_left = False
        folded_ix = []
        for t, ml in enumerate(multilinears):
            if isinstance(ml, TransparentMultilinear):
                if ml.switchover_round == 0:
                    if tensor_query is None:
                        raise InputValidation("tensor_query is required while a multilinear is transparent")
                    d = self._partial_eval(evaluation_order, ml, tensor_query)
                    multilinears[t] = FoldedMultilinear(d, ml.const_suffix[0] if ml.const_suffix[1] else 0)
                else:
                    ml.switchover_round -= 1
                    any_transparent_left = True
            else:
                folded_ix.append(t)
        m = len(folded_ix)
        if m == 0:
            return any_transparent_left
        mls = [multilinears[t] for t in folded_ix]
        ptrs = (C.c_void_p * m)(*[ml.evals.ptr for ml in mls])
        prefix = (C.c_uint64 * m)(*[min(ml.evals.len(), 1 << n_vars) for ml in mls])
        new_lens = (C.c_uint64 * m)()
        sfx = _u64_list([ml.suffix_eval for ml in mls])
        if evaluation_order == EvaluationOrder.HighToLow:
            L._check(L._lib.b200_fold_multilinears_high_to_low(L._ctx, ptrs, m, n_vars, prefix, sfx, _u64x2(challenge), new_lens))
            for ml, n in zip(mls, new_lens):
                ml.evals = ml.evals.slice(0, int(n))
        else:
            outs = [self._take((int(p) + 1) // 2) for p in prefix]
            optrs = (C.c_void_p * m)(*[o.ptr for o in outs])
            L._check(L._lib.b200_fold_multilinears_low_to_high(L._ctx, ptrs, optrs, m, n_vars, prefix, sfx, _u64x2(challenge), new_lens))
            for ml, o, n in zip(mls, outs, new_lens):
                self._give(ml.evals)
                ml.evals = o.slice(0, int(n))
        return any_transparent_left

    # -- prove/common.rs:13-73 fold_partial_eq_ind, HighToLow: E'[i] = E[i] + E[half + i]
    def fold_partial_eq_ind(self, n_vars: int, eq_ind: DevSlice, evaluation_order: EvaluationOrder = EvaluationOrder.HighToLow) -> DevSlice:
        if n_vars == 0:
            return eq_ind
        if self._tail is not None and self._tail["eq_ptr"] == eq_ind.ptr and self._tail["eq_pending"] > 0:
            self._tail["eq_pending"] -= 1  # halved in place by the tail kernel
            return eq_ind.slice(0, 1 << (n_vars - 1))
        if evaluation_order == EvaluationOrder.LowToHigh:
            # E'[i] = E[2i] + E[2i+1] (common.rs:50-58) = fold_right of E by the all-ones pair
            if self._ones2 is None:
                self._ones2 = self._l.dev_alloc(2)
                self._l.fill(self._ones2, 1)
            out = self._take(1 << (n_vars - 1))
            self._l._check(self._l._lib.b200_fold_right(self._l._ctx, eq_ind.ptr, eq_ind.len(), 7, self._ones2.ptr, 2, out.ptr, out.len()))
            self._give(eq_ind)
            return out
        lo, hi = eq_ind.split_half()
        self._l._check(self._l._lib.b200_kernel_add(self._l._ctx, n_vars - 1, lo.ptr, hi.ptr, lo.ptr))
        return lo

    # -- backend.rs:78-82: partial evaluation of a B128 multilinear on its HIGH variables by an
    #    expanded tensor query (fold_left, math/src/fold.rs:88-179)
    def evaluate_partial_high(self, multilinear: DevSlice, query_expansion: DevSlice) -> DevSlice:
        n, q = multilinear.len(), query_expansion.len()
        if q == 0 or n % q:
            raise InputValidation("query expansion must divide the multilinear")
        out = self._l.dev_alloc(n // q)
        self._l._check(self._l._lib.b200_fold_left(self._l._ctx, multilinear.ptr, n, 7, query_expansion.ptr, q, out.ptr, out.len()))
        return out
