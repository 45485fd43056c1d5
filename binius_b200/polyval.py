"""BinaryField128bPolyval -- the reference's "fast" field (crates/field/src/polyval.rs) -- and the GKR grand-product
data plane (crates/core/src/protocols/gkr_gpa) on the tower kernels.

The reference runs the grand-product argument in POLYVAL because CLMUL makes that field cheap on a CPU
(`convert_witnesses_to_fast_ext`, core/src/constraint_system/prove.rs:291-292).  POLYVAL is ISOMORPHIC to
BinaryField128b: phi(a * b) = phi(a) * phi(b) on the stored representations, phi = the GF(2)-linear basis change of
polyval.rs:516-788.  A B200 has no carry-less multiplier, so here POLYVAL data is mapped to the tower basis once
(`b200_linear_map`, the byte-LUT kernel), every layer product and every sumcheck round runs on the tower kernels
(tensor-core inner products, LUT folds), and the handful of scalars that cross the interface per round is mapped
back -- bit-identical to the Montgomery arithmetic (tests/test_gpu_polyval.py).  The basis change is not copied from
the reference: the library derives it from the two published multiplicative generators
(`b200_host_polyval_basis_change`) and tests compare it with the reference's tables.
"""
from __future__ import annotations

import ctypes as C
from functools import lru_cache
from typing import List, Sequence

import numpy as np

from . import _lib
from .layer import ArithCircuit, B200Layer, DevSlice, InputValidation, SlicesBatch, _u64x2

ONE = 0xC2000000000000000000000000000001  # polyval.rs:262 (X^128 mod p: the Montgomery form of 1)
_M64 = (1 << 64) - 1


@lru_cache(maxsize=1)
def basis_change():
    """(tower_to_polyval, polyval_to_tower): 128 basis images each, python ints"""
    lib = _lib.load()
    t2p, p2t = (C.c_uint64 * 256)(), (C.c_uint64 * 256)()
    if lib.b200_host_polyval_basis_change(t2p, p2t):
        raise RuntimeError("basis change derivation failed")
    f = lambda a: [int(a[2 * k]) | (int(a[2 * k + 1]) << 64) for k in range(128)]  # noqa: E731
    return f(t2p), f(p2t)


def _lin(images: Sequence[int], x: int) -> int:
    acc, k = 0, 0
    while x:
        if x & 1:
            acc ^= images[k]
        x >>= 1
        k += 1
    return acc


def to_polyval(x: int) -> int:
    """From<BinaryField128b> for BinaryField128bPolyval (polyval.rs:648-652)"""
    return _lin(basis_change()[0], x)


def to_tower(x: int) -> int:
    """From<BinaryField128bPolyval> for BinaryField128b (polyval.rs:790-794)"""
    return _lin(basis_change()[1], x)


def mul(a: int, b: int) -> int:
    """montgomery_multiply on stored forms (arch/portable/packed_polyval_128.rs:88-122), host scalar"""
    A2 = C.c_uint64 * 2
    out = A2()
    _lib.load().b200_host_polyval_mul(A2(a & _M64, a >> 64), A2(b & _M64, b >> 64), out)
    return int(out[0]) | (int(out[1]) << 64)


def _images_arg(images: Sequence[int]):
    arr = (C.c_uint64 * 256)()
    for k, v in enumerate(images):
        arr[2 * k], arr[2 * k + 1] = v & _M64, v >> 64
    return arr


def linear_map(hal: B200Layer, src: DevSlice, dst: DevSlice, images: Sequence[int]):
    """FieldLinearTransformation::transform on a device slice (in place when dst is src)"""
    if src.len() != dst.len():
        raise InputValidation("src and dst must have the same length")
    hal._check(hal._lib.b200_linear_map(hal._ctx, src.ptr, dst.ptr, src.len(), _images_arg(images)))


def convert_to_tower(hal: B200Layer, data: DevSlice, out: DevSlice = None) -> DevSlice:
    """POLYVAL-represented device data -> tower basis (what the device computes in)"""
    out = out or data
    linear_map(hal, data, out, basis_change()[1])
    return out


def convert_to_fast_ext(hal: B200Layer, data: DevSlice, out: DevSlice = None) -> DevSlice:
    """convert_witnesses_to_fast_ext (prove.rs:291-292) on the device: tower -> POLYVAL representation"""
    out = out or data
    linear_map(hal, data, out, basis_change()[0])
    return out


class GrandProductWitness:
    """GrandProductWitness::new (gkr_gpa/gkr_gpa.rs:40-90): n_vars + 1 layers, layer k+1 = low half * high half of
    layer k.  The input is POLYVAL-represented (as the prover hands it over) unless `input_is_tower`; the layers live on
    the device in the TOWER basis; accessors map back."""

    def __init__(self, hal: B200Layer, n_vars: int, input_layer: DevSlice, input_is_tower: bool = False):
        if input_layer.len() != 1 << n_vars:
            raise InputValidation("NumberOfVariablesMismatch")  # (truncated witnesses are padded with ONE by the caller)
        self.hal, self.n_vars = hal, n_vars
        store = hal.dev_alloc(max((2 << n_vars) - 1, 1))
        first = store.slice(0, 1 << n_vars)
        if input_is_tower:
            hal.copy_d2d(input_layer, first)
        else:
            convert_to_tower(hal, input_layer, first)
        self.layers: List[DevSlice] = [first]
        prod = hal.compile_expr(ArithCircuit.var(0) * ArithCircuit.var(1))
        off = 1 << n_vars

        def build(ex):
            nonlocal off
            for k in range(n_vars):
                prev = self.layers[-1]
                half = prev.len() // 2
                cur = store.slice(off, off + half)
                ex.compute_composite(SlicesBatch([prev.slice(0, half), prev.slice(half, 2 * half)], half), cur, prod)
                self.layers.append(cur)
                off += half
            return []

        hal.execute(build)

    def layer_polyval(self, k: int) -> np.ndarray:
        """layer k in the POLYVAL representation, on the host ((n, 2) uint64)"""
        tmp = self.hal.dev_alloc(self.layers[k].len())
        convert_to_fast_ext(self.hal, self.layers[k], tmp)
        out = self.hal.to_host(tmp)
        self.hal.dev_free(tmp)
        return out

    def grand_product_evaluation(self) -> int:
        v = self.hal.to_host(self.layers[-1])
        return to_polyval(int(v[0, 0]) | (int(v[0, 1]) << 64))


def gpa_round_evals(backend, n_vars: int, layer: DevSlice, eq_ind: DevSlice, first_round_eval_1s: bool = False) -> List[int]:
    """One round of the GPA layer sumcheck (gkr_gpa/prove.rs: eq-ind sumcheck of the product of the two half-layer
    multilinears): `layer` (tower basis, 2^(n_vars+1) elements) is split into its halves A | B; returns the round values
    [at 1 (unless known), at infinity] in the POLYVAL representation."""
    from .hal import EqIndEvaluator, FoldedMultilinear

    half = layer.len() // 2
    if half != 1 << n_vars:
        raise InputValidation("layer must hold 2^(n_vars + 1) elements")
    mls = [FoldedMultilinear(layer.slice(0, half), 0), FoldedMultilinear(layer.slice(half, 2 * half), 0)]
    ev = EqIndEvaluator(ArithCircuit.var(0) * ArithCircuit.var(1), have_first_round_eval_1s=first_round_eval_1s)
    got = backend.sumcheck_compute_round_evals(n_vars, mls, [ev], eq_ind, [])
    return [to_polyval(v) for v in got[0]]
