"""Host-side mirror of the old HAL `binius_hal::ComputationBackend` (reference
crates/hal/src/backend.rs:35-83) for device-resident multilinears, over the C ABI.

    tensor_product_full_query     backend.rs:42-45  -> math/src/tensor_prod_eq_ind.rs:94-101
    sumcheck_compute_round_evals  backend.rs:48-62  -> hal/src/sumcheck_round_calculation.rs:126-349
    sumcheck_fold_multilinears    backend.rs:65-75  -> hal/src/sumcheck_folding.rs:149-241
    evaluate_partial_high         backend.rs:78-82  -> math/src/multilinear_extension.rs:253-300

The reference's trait is generic over packed fields, multilinear trait objects and evaluator trait
objects; this backend covers what the prover's hot loop instantiates it with after the switchover:
`SumcheckMultilinear::Folded` B128 multilinears (hal/src/sumcheck_multilinear.rs:8-40), `ArithCircuit`
compositions, HighToLow evaluation order, and the eq-ind evaluator of
core/src/protocols/sumcheck/prove/eq_ind.rs:646-731.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence

from .layer import (ArithCircuit, B200Layer, DevSlice, ExprEval, InputValidation, _u64_list, _u64x2)


@dataclass
class FoldedMultilinear:
    """SumcheckMultilinear::Folded { large_field_folded_evals, suffix_eval }: `evals` holds the stored
    prefix (its length may be shorter than 2^n_vars); the rest of the hypercube equals suffix_eval."""
    evals: DevSlice
    suffix_eval: int = 0


@dataclass
class EqIndEvaluator:
    """The data of the eq-ind `Evaluator` (eq_ind.rs:646-662) that reaches the backend: the
    composition, and whether r(1) is already known (`have_first_round_eval_1s`)."""
    composition: ArithCircuit
    have_first_round_eval_1s: bool = False

    def degree(self) -> int:
        cached = getattr(self.composition, "_degree", None)
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

        d = deg(self.composition, len(self.composition.steps) - 1)
        self.composition._degree = d
        return d

    def eval_point_indices(self) -> range:
        # eq_ind.rs:667-671: skip r(1) in the first round when known; never evaluate at 0
        return range(2 if self.have_first_round_eval_1s else 1, self.degree() + 1)


class B200Backend:
    """ComputationBackend over one B200Layer."""

    def __init__(self, layer: B200Layer):
        self._l = layer
        self._exprs = {}

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

    # -- backend.rs:48-62
    def sumcheck_compute_round_evals(self, n_vars: int, multilinears: Sequence[FoldedMultilinear],
                                     evaluators: Sequence[EqIndEvaluator], eq_ind_partial_evals: DevSlice,
                                     finite_evaluation_points: Sequence[int]) -> List[List[int]]:
        """Returns RoundEvals per evaluator: the values at eval_point_indices() (1 = at 1, 2 = infinity,
        k >= 3 = finite_evaluation_points[k-3]); sumcheck_round_calculation.rs:156-163 length check."""
        if n_vars == 0:
            raise InputValidation("Computing round evaluations requires at least a single variable.")
        hi = max([ev.eval_point_indices().stop for ev in evaluators], default=0)
        if len(finite_evaluation_points) != max(hi - 3, 0):
            raise InputValidation("IncorrectNontrivialEvalPointsLength")
        if eq_ind_partial_evals.len() != 1 << (n_vars - 1):
            raise InputValidation("eq_ind_partial_evals must have 2^(n_vars-1) elements")
        lo = min([ev.eval_point_indices().start for ev in evaluators], default=1)
        codes = list(range(lo, hi))
        if not codes or not evaluators:
            return [[] for _ in evaluators]
        pts = [0 if c < 3 else finite_evaluation_points[c - 3] for c in codes]
        m = len(multilinears)
        ptrs = (C.c_void_p * max(m, 1))(*[ml.evals.ptr for ml in multilinears])
        lens = (C.c_uint64 * max(m, 1))(*[ml.evals.len() for ml in multilinears])
        sfx = _u64_list([ml.suffix_eval for ml in multilinears])
        compiled = [self._compiled(ev.composition) for ev in evaluators]
        comps = (C.c_void_p * len(evaluators))(*[c[0].handle.value for c in compiled])
        leads = (C.c_void_p * len(evaluators))(*[c[1].handle.value for c in compiled])
        ccodes = (C.c_uint32 * len(codes))(*codes)
        first = C.c_uint32()
        L = self._l
        L._check(L._lib.b200_results_reset(L._ctx))
        L._check(L._lib.b200_eq_ind_round_evals(L._ctx, ptrs, lens, sfx, m, n_vars, eq_ind_partial_evals.ptr, comps, leads,
                                                len(evaluators), ccodes, _u64_list(pts), len(codes), C.byref(first)))
        total = len(evaluators) * len(codes)
        slots = (C.c_uint32 * total)(*range(first.value, first.value + total))
        out = (C.c_uint64 * (2 * total))()
        L._check(L._lib.b200_results_fetch(L._ctx, slots, total, out))
        vals = [int(out[2 * i]) | (int(out[2 * i + 1]) << 64) for i in range(total)]
        res = []
        for e, ev in enumerate(evaluators):
            rng = ev.eval_point_indices()
            res.append([vals[e * len(codes) + (k - lo)] for k in rng])
        return res

    # -- backend.rs:65-75 (HighToLow, Folded branch): folds in place, truncates the handles
    def sumcheck_fold_multilinears(self, n_vars: int, multilinears: List[FoldedMultilinear], challenge: int) -> bool:
        m = len(multilinears)
        if m == 0:
            return False
        ptrs = (C.c_void_p * m)(*[ml.evals.ptr for ml in multilinears])
        prefix = (C.c_uint64 * m)(*[min(ml.evals.len(), 1 << n_vars) for ml in multilinears])
        new_lens = (C.c_uint64 * m)()
        L = self._l
        L._check(L._lib.b200_fold_multilinears_high_to_low(L._ctx, ptrs, m, n_vars, prefix,
                                                           _u64_list([ml.suffix_eval for ml in multilinears]), _u64x2(challenge), new_lens))
        for ml, n in zip(multilinears, new_lens):
            ml.evals = ml.evals.slice(0, int(n))
        return False  # no transparent multilinears left (sumcheck_folding.rs:245-262)

    # -- prove/common.rs:13-73 fold_partial_eq_ind, HighToLow: E'[i] = E[i] + E[half + i]
    def fold_partial_eq_ind(self, n_vars: int, eq_ind: DevSlice) -> DevSlice:
        if n_vars == 0:
            return eq_ind
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
