"""Ring-switch / evalcheck host paths that are O(trace) passes over committed data, on the device:

  evaluate_partial_high   MultilinearExtension::evaluate_partial_high (crates/math/src/multilinear_extension.rs:253-293)
                          of a packed sub-field multilinear by an expanded tensor query = fold_left
  compute_partial_evals   crates/core/src/ring_switch/prove.rs:147-208: for every PIOP sumcheck claim the committed
                          multilinear is partially evaluated on its high variables by the claim's suffix query and the
                          first 2^kappa values form the claim's tensor-algebra element (repeated cyclically when the
                          partial evaluation is shorter); suffix queries are expanded once per distinct suffix
                          (MemoizedData::memoize_query_par)

For bit-packed (B1) witnesses and kappa = 7 the fold is an outer-product bit-GEMM on the tensor cores
(b200_fold_left's long-query fast path, roundevals_tc.cuh); other levels take the generic kernels.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

from .layer import B200Layer, DevSlice, InputValidation, to_ints


@dataclass
class CommittedWitness:
    """a committed multilinear in MLEEmbeddingAdapter layout: 2^n_vars scalars of tower level `tower_level`
    packed into B128 words"""
    evals: DevSlice
    tower_level: int
    n_vars: int


def evaluate_partial_high(hal: B200Layer, w: CommittedWitness, query_expansion: DevSlice) -> DevSlice:
    n_scalars, q = 1 << w.n_vars, query_expansion.len()
    if q == 0 or n_scalars % q:
        raise InputValidation("query expansion must divide the multilinear")
    out = hal.dev_alloc(n_scalars // q)
    hal._check(hal._lib.b200_fold_left(hal._ctx, w.evals.ptr, w.evals.len(), w.tower_level, query_expansion.ptr, q, out.ptr, out.len()))
    return out


def compute_partial_evals(backend, witnesses: Sequence[CommittedWitness],
                          claims: Sequence[Tuple[int, Tuple[int, ...], int]]) -> List[List[int]]:
    """claims: (committed_idx, suffix point, kappa) per PIOP sumcheck claim -> the 2^kappa elements of each claim's
    TowerTensorAlgebra (ring_switch/prove.rs:171-203)"""
    hal = backend._l
    memo: Dict[Tuple[int, ...], DevSlice] = {}
    for _, suffix, _ in claims:
        if tuple(suffix) not in memo:
            memo[tuple(suffix)] = backend.tensor_product_full_query(list(suffix))
    out = []
    for committed_idx, suffix, kappa in claims:
        pe = evaluate_partial_high(hal, witnesses[committed_idx], memo[tuple(suffix)])
        vals = to_ints(hal.to_host(pe))[: 1 << kappa]
        hal.dev_free(pe)
        while len(vals) < (1 << kappa):
            vals = (vals * 2)[: 1 << kappa]
        out.append(vals)
    return out
