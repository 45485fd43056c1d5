"""Host-side mirror of the reference's `binius_compute` trait surface over the C ABI.

Names, argument meaning and error behaviour follow the reference so that the parity tests read like
`crates/compute_test_utils/src/layer.rs`:

    ComputeLayer            crates/compute/src/layer.rs:22-88      -> B200Layer
    ComputeLayerExecutor    crates/compute/src/layer.rs:100-510    -> B200Executor
    KernelExecutor          crates/compute/src/layer.rs:518-590    -> B200KernelExecutor
    KernelMemMap/Buffer     crates/compute/src/layer.rs:595-704    -> KernelMemMap / KernelBuffer
    ComputeMemory           crates/compute/src/memory.rs:69-235    -> DevSlice (ALIGNMENT = 1)
    SubfieldSlice           crates/compute/src/memory.rs:257-281   -> SubfieldSlice
    BumpAllocator           crates/compute/src/alloc.rs:31-105     -> BumpAllocator / HostBumpAllocator
    ComputeHolder/Data      crates/compute/src/layer.rs:732-776    -> B200LayerHolder / ComputeData
    ArithCircuit            crates/math/src/arith_expr.rs:200-383  -> ArithCircuit

Field elements are python ints < 2^128 (scalars) or numpy uint64 arrays of shape (n, 2) = [lo, hi]
(host vectors).  All compute runs in the sm_100a library; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _lib

TOWER_LEVEL = 7  # F = BinaryField128b


# ------------------------------------------------------------------------------------------------ errors
class Error(Exception):
    """compute::Error (crates/compute/src/layer.rs:706-716)"""


class InputValidation(Error):
    pass


class AllocError(Error):
    """alloc::Error::OutOfMemory (crates/compute/src/alloc.rs:110-114)"""


class DeviceError(Error):
    pass


class NttError(Error):
    """ntt::Error (crates/ntt/src/error.rs:3-29); `.kind` names the variant."""
    KINDS = {11: "PowerOfTwoLengthRequired", 12: "SkipRoundsTooLarge", 13: "BatchTooLarge",
             14: "CosetIndexOutOfBounds", 15: "DomainTooSmall", 16: "FieldTooSmall"}

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code
        self.kind = self.KINDS.get(code, "Unknown")


def _raise(code: int, msg: str):
    if code == _lib.ERR_INPUT_VALIDATION:
        raise InputValidation(msg)
    if code == _lib.ERR_ALLOC:
        raise AllocError(msg)
    if code in NttError.KINDS:
        raise NttError(code, msg)
    raise DeviceError(msg or f"device error (status {code})")


def _u64x2(v: int):
    return (C.c_uint64 * 2)(v & 0xFFFFFFFFFFFFFFFF, (v >> 64) & 0xFFFFFFFFFFFFFFFF)


def _u64_list(vals: Sequence[int]):
    arr = (C.c_uint64 * max(2 * len(vals), 2))()
    for i, v in enumerate(vals):
        arr[2 * i] = v & 0xFFFFFFFFFFFFFFFF
        arr[2 * i + 1] = (v >> 64) & 0xFFFFFFFFFFFFFFFF
    return arr


def to_arr(vals: Sequence[int]) -> np.ndarray:
    a = np.empty((len(vals), 2), dtype=np.uint64)
    for i, v in enumerate(vals):
        a[i, 0] = v & 0xFFFFFFFFFFFFFFFF
        a[i, 1] = v >> 64
    return a


def to_ints(arr) -> List[int]:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 2)
    return [int(lo) | (int(hi) << 64) for lo, hi in arr]


# ------------------------------------------------------------------------------------------------ memory
class DevSlice:
    """Opaque handle to a slice of F elements in device memory (FSlice / FSliceMut).
    ComputeMemory::ALIGNMENT = 1 (crates/compute/src/memory.rs:69-72): any element offset is valid."""
    ALIGNMENT = 1
    __slots__ = ("ptr", "_len", "mutable")

    def __init__(self, ptr: int, length: int, mutable: bool = True):
        self.ptr, self._len, self.mutable = ptr, length, mutable

    def __len__(self):
        return self._len

    def len(self):
        return self._len

    def is_empty(self):
        return self._len == 0

    def __repr__(self):
        return f"DevSlice(0x{self.ptr:x}, len={self._len}, {'mut' if self.mutable else 'const'})"

    # ComputeMemory::{as_const, to_const, narrow, narrow_mut, to_owned_mut}
    def as_const(self) -> "DevSlice":
        return DevSlice(self.ptr, self._len, False)

    to_const = as_const

    def narrow(self) -> "DevSlice":
        return DevSlice(self.ptr, self._len, self.mutable)

    # ComputeMemory::slice / slice_mut
    def slice(self, start: int = 0, end: Optional[int] = None) -> "DevSlice":
        end = self._len if end is None else end
        if not (0 <= start <= end <= self._len):
            raise IndexError(f"range {start}..{end} out of bounds for slice of length {self._len}")
        return DevSlice(self.ptr + 16 * start, end - start, self.mutable)

    slice_mut = slice

    def __getitem__(self, key):
        if isinstance(key, slice):
            s, e, st = key.indices(self._len)
            assert st == 1
            return self.slice(s, e)
        raise TypeError("DevSlice supports only range indexing; use ComputeLayer.copy_d2h to read elements")

    # ComputeMemory::{split_at, split_at_mut, split_half, split_half_mut, slice_chunks}
    def split_at(self, mid: int):
        return self.slice(0, mid), self.slice(mid, self._len)

    split_at_mut = split_at

    def split_half(self):
        if self._len % 2:
            raise AssertionError("split_half requires an even length")
        return self.split_at(self._len // 2)

    split_half_mut = split_half

    def slice_chunks(self, chunk_len: int):
        if chunk_len == 0 or self._len % chunk_len:
            raise AssertionError("chunk_len must divide the slice length")
        return [self.slice(i, i + chunk_len) for i in range(0, self._len, chunk_len)]

    slice_chunks_mut = slice_chunks


@dataclass
class SubfieldSlice:
    """memory.rs:257-281: `slice` viewed as packed sub-field elements of tower level `tower_level`."""
    slice: DevSlice
    tower_level: int

    def len(self):
        return self.slice.len() << (TOWER_LEVEL - self.tower_level)

    def is_empty(self):
        return self.slice.is_empty()


class SlicesBatch:
    """memory.rs:29-66"""

    def __init__(self, rows: Sequence[DevSlice], row_len: int):
        for r in rows:
            assert r.len() == row_len, "all rows must have length row_len"
        self.rows, self._row_len = list(rows), row_len

    def n_rows(self):
        return len(self.rows)

    def row_len(self):
        return self._row_len

    def row(self, i):
        return self.rows[i]

    def iter(self):
        return iter(self.rows)


class BumpAllocator:
    """alloc.rs:31-105 over device memory"""

    def __init__(self, buffer: DevSlice):
        self._buf, self._off = buffer, 0

    def alloc(self, n: int) -> DevSlice:
        if n > self._buf.len() - self._off:
            raise AllocError("allocator is out of memory")
        s = self._buf.slice(self._off, self._off + n)
        self._off += n
        return s

    def remaining(self):
        return self._buf.len() - self._off

    def capacity(self):
        return self._buf.len()

    def subscope_allocator(self) -> "BumpAllocator":
        return BumpAllocator(self._buf.slice(self._off, self._buf.len()))


class HostBumpAllocator:
    """alloc.rs:116-180 over a host numpy arena"""

    def __init__(self, buffer: np.ndarray):
        self._buf, self._off = buffer, 0

    def alloc(self, n: int) -> np.ndarray:
        if n > len(self._buf) - self._off:
            raise AllocError("allocator is out of memory")
        s = self._buf[self._off:self._off + n]
        self._off += n
        return s

    def remaining(self):
        return len(self._buf) - self._off

    def capacity(self):
        return len(self._buf)


# ------------------------------------------------------------------------------------------------ expressions
class ArithCircuit:
    """math/src/arith_expr.rs:200-383; steps are ('add',l,r)|('mul',l,r)|('pow',l,e)|('const',c)|('var',i)."""

    def __init__(self, steps):
        self.steps = list(steps)

    @staticmethod
    def var(i):
        return ArithCircuit([("var", i)])

    @staticmethod
    def constant(c):
        return ArithCircuit([("const", c)])

    @staticmethod
    def zero():
        return ArithCircuit.constant(0)

    @staticmethod
    def one():
        return ArithCircuit.constant(1)

    def _shifted(self, off):
        out = []
        for st in self.steps:
            if st[0] in ("add", "mul"):
                out.append((st[0], st[1] + off, st[2] + off))
            elif st[0] == "pow":
                out.append(("pow", st[1] + off, st[2]))
            else:
                out.append(st)
        return out

    def _binop(self, other, op):
        a = list(self.steps)
        b = other._shifted(len(a))
        return ArithCircuit(a + b + [(op, len(a) - 1, len(a) + len(b) - 1)])

    def __add__(self, o):
        return self._binop(o, "add")

    __sub__ = __add__

    def __mul__(self, o):
        return self._binop(o, "mul")

    def pow(self, e):
        return ArithCircuit(self.steps + [("pow", len(self.steps) - 1, e)])

    def n_vars(self):
        return 1 + max([s[1] for s in self.steps if s[0] == "var"], default=-1)

    def _lt(self, step):
        st = self.steps[step]
        if st[0] == "const":
            return 0, ArithCircuit.constant(st[1])
        if st[0] == "var":
            return 1, ArithCircuit.var(st[1])
        if st[0] == "add":
            dl, l = self._lt(st[1])
            dr, r = self._lt(st[2])
            if dl < dr:
                return dr, r
            if dl > dr:
                return dl, l
            return dl, l + r
        if st[0] == "mul":
            dl, l = self._lt(st[1])
            dr, r = self._lt(st[2])
            return dl + dr, l * r
        db, b = self._lt(st[1])
        return db * st[2], b.pow(st[2])

    def leading_term(self) -> "ArithCircuit":
        """arith_expr.rs:334-365"""
        return self._lt(len(self.steps) - 1)[1]

    def encode(self):
        arr = (_lib.ExprStep * max(len(self.steps), 1))()
        for i, st in enumerate(self.steps):
            k = st[0]
            if k == "add":
                arr[i] = _lib.ExprStep(0, st[1], st[2], 0, 0)
            elif k == "mul":
                arr[i] = _lib.ExprStep(1, st[1], st[2], 0, 0)
            elif k == "pow":
                arr[i] = _lib.ExprStep(2, st[1], st[2], 0, 0)
            elif k == "const":
                arr[i] = _lib.ExprStep(3, 0, 0, st[1] & 0xFFFFFFFFFFFFFFFF, st[1] >> 64)
            elif k == "var":
                arr[i] = _lib.ExprStep(4, st[1], 0, 0, 0)
            else:
                raise ValueError(k)
        return arr


class ExprEval:
    """ComputeLayerExecutor::ExprEval: a compiled ArithCircuit living on the device."""

    def __init__(self, layer: "B200Layer", circuit: ArithCircuit):
        self.circuit = circuit
        self._lib = layer._lib
        h = C.c_void_p()
        layer._check(self._lib.b200_expr_compile(layer._ctx, circuit.encode(), len(circuit.steps), C.byref(h)))
        self.handle = h

    def n_vars(self):
        return self._lib.b200_expr_n_vars(self.handle)

    def __del__(self):
        try:
            self._lib.b200_expr_free(self.handle)
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------ kernels
class KernelMemMap:
    """layer.rs:595-676"""

    def __init__(self, kind, data=None, log_min_chunk_size=0, log_size=0):
        self.kind, self.data, self.log_min_chunk_size, self.log_size = kind, data, log_min_chunk_size, log_size

    @staticmethod
    def Chunked(data: DevSlice, log_min_chunk_size: int):
        return KernelMemMap("chunked", data.as_const(), log_min_chunk_size)

    @staticmethod
    def ChunkedMut(data: DevSlice, log_min_chunk_size: int):
        return KernelMemMap("chunked_mut", data, log_min_chunk_size)

    @staticmethod
    def Local(log_size: int):
        return KernelMemMap("local", None, 0, log_size)

    @staticmethod
    def log_chunks_range(mappings):
        lo, hi = None, None
        for m in mappings:
            if m.kind == "local":
                r = (0, m.log_size)
            else:
                n = m.data.len()
                if n == 0 or n & (n - 1):
                    raise InputValidation("kernel buffers must have power-of-two length")
                log_n = n.bit_length() - 1
                mn = min(max(m.log_min_chunk_size, 0), log_n)  # log2(ALIGNMENT) = 0
                r = (0, log_n - mn)
            lo = r[0] if lo is None else max(lo, r[0])
            hi = r[1] if hi is None else min(hi, r[1])
        return None if lo is None else (lo, hi)


class KernelBuffer:
    """layer.rs:678-704"""

    def __init__(self, data: DevSlice, mutable: bool):
        self.data, self.mutable = data, mutable

    def to_ref(self) -> DevSlice:
        return self.data.as_const()

    def len(self):
        return self.data.len()


class OpValue:
    """Deferred scalar (ComputeLayerExecutor::OpValue / KernelExecutor::Value): a result slot."""
    __slots__ = ("slot",)

    def __init__(self, slot):
        self.slot = slot


class B200KernelExecutor:
    """KernelExecutor (layer.rs:518-590).  The layer always selects log_chunks = 0, so every kernel
    op is a grid-wide launch over the whole mapped buffers (allowed by layer.rs:149-160)."""

    def __init__(self, layer: "B200Layer"):
        self._l = layer

    def decl_value(self, init: int) -> OpValue:
        s = C.c_uint32()
        self._l._check(self._l._lib.b200_kernel_decl_value(self._l._ctx, _u64x2(init), C.byref(s)))
        return OpValue(s.value)

    def sum_composition_evals(self, inputs: SlicesBatch, composition: ExprEval, batch_coeff: int, accumulator: OpValue):
        ptrs = (C.c_void_p * max(inputs.n_rows(), 1))(*[r.ptr for r in inputs.rows])
        self._l._check(self._l._lib.b200_kernel_sum_composition_evals(
            self._l._ctx, ptrs, inputs.n_rows(), inputs.row_len(), composition.handle, _u64x2(batch_coeff), accumulator.slot))

    def add(self, log_len: int, src1: DevSlice, src2: DevSlice, dst: DevSlice):
        assert src1.len() == src2.len() == dst.len() == 1 << log_len
        self._l._check(self._l._lib.b200_kernel_add(self._l._ctx, log_len, src1.ptr, src2.ptr, dst.ptr))

    def add_assign(self, log_len: int, src: DevSlice, dst: DevSlice):
        assert src.len() == dst.len() == 1 << log_len
        self._l._check(self._l._lib.b200_kernel_add_assign(self._l._ctx, log_len, src.ptr, dst.ptr))


class B200Executor:
    """ComputeLayerExecutor (layer.rs:100-510)"""

    def __init__(self, layer: "B200Layer"):
        self._l = layer
        self._lib = layer._lib
        self._ctx = layer._ctx
        self._local = []

    # -- combinators (layer.rs:115-132): one in-order stream, so sequential issue is a valid schedule
    def join(self, op1: Callable, op2: Callable):
        return op1(self), op2(self)

    def map(self, items, f: Callable):
        return [f(self, it) for it in items]

    # -- kernels
    def _map_kernel_mem(self, mem_maps):
        rng = KernelMemMap.log_chunks_range(mem_maps)
        if rng is None:
            raise InputValidation("Many variant must have at least one entry")
        bufs = []
        for m in mem_maps:
            if m.kind == "local":
                # KernelMemMap::Local: library-owned scratch of the open scope (never written when the scope fuses)
                p = C.c_void_p()
                self._l._check(self._lib.b200_kernel_local(self._ctx, m.log_size, C.byref(p)))
                bufs.append(KernelBuffer(DevSlice(p.value, 1 << m.log_size), True))
            else:
                bufs.append(KernelBuffer(m.data, m.kind == "chunked_mut"))
        return bufs

    def _kernel_scope(self, map_fn, mem_maps):
        self._l._check(self._lib.b200_kernel_scope_begin(self._ctx))
        try:
            bufs = self._map_kernel_mem(mem_maps)
            out = map_fn(B200KernelExecutor(self._l), 0, bufs)
        except BaseException:
            self._lib.b200_kernel_scope_end(self._ctx)  # release the scope; the original error wins
            raise
        self._l._check(self._lib.b200_kernel_scope_end(self._ctx))
        return out

    def accumulate_kernels(self, map_fn: Callable, mem_maps: List[KernelMemMap]) -> List[OpValue]:
        """layer.rs:134-192; closure signature map_fn(kernel_exec, log_chunks, buffers) -> [Value].  The ops the
        closure issues are recorded by the library and lowered together when the scope ends."""
        return list(self._kernel_scope(map_fn, mem_maps))

    def map_kernels(self, map_fn: Callable, mem_maps: List[KernelMemMap]):
        """layer.rs:194-245"""
        self._kernel_scope(map_fn, mem_maps)

    # -- ops
    def inner_product(self, a_in: SubfieldSlice, b_in: DevSlice) -> OpValue:
        s = C.c_uint32()
        self._l._check(self._lib.b200_inner_product(self._ctx, a_in.slice.ptr, a_in.slice.len(), a_in.tower_level,
                                                    b_in.ptr, b_in.len(), C.byref(s)))
        return OpValue(s.value)

    def tensor_expand(self, log_n: int, coordinates: Sequence[int], data: DevSlice):
        self._l._check(self._lib.b200_tensor_expand(self._ctx, data.ptr, data.len(), log_n, _u64_list(coordinates), len(coordinates)))

    def fold_left(self, mat: SubfieldSlice, vec: DevSlice, out: DevSlice):
        self._l._check(self._lib.b200_fold_left(self._ctx, mat.slice.ptr, mat.slice.len(), mat.tower_level, vec.ptr, vec.len(), out.ptr, out.len()))

    def fold_right(self, mat: SubfieldSlice, vec: DevSlice, out: DevSlice):
        self._l._check(self._lib.b200_fold_right(self._ctx, mat.slice.ptr, mat.slice.len(), mat.tower_level, vec.ptr, vec.len(), out.ptr, out.len()))

    def fri_fold(self, ntt, log_len: int, log_batch_size: int, challenges: Sequence[int], data_in: DevSlice, data_out: DevSlice):
        self._l._check(self._lib.b200_fri_fold(self._ctx, ntt.handle, log_len, log_batch_size, _u64_list(challenges), len(challenges),
                                               data_in.ptr, data_in.len(), data_out.ptr, data_out.len()))

    def extrapolate_line(self, evals_0: DevSlice, evals_1: DevSlice, z: int):
        self._l._check(self._lib.b200_extrapolate_line(self._ctx, evals_0.ptr, evals_0.len(), evals_1.ptr, evals_1.len(), _u64x2(z)))

    def compute_composite(self, inputs: SlicesBatch, output: DevSlice, composition: ExprEval):
        if composition.circuit.n_vars() != inputs.n_rows():
            raise InputValidation("composition not match with input")
        ptrs = (C.c_void_p * max(inputs.n_rows(), 1))(*[r.ptr for r in inputs.rows])
        self._l._check(self._lib.b200_compute_composite(self._ctx, ptrs, inputs.n_rows(), inputs.row_len(), output.ptr, output.len(), composition.handle))

    def pairwise_product_reduce(self, input: DevSlice, round_outputs: List[DevSlice]):
        n = len(round_outputs)
        ptrs = (C.c_void_p * max(n, 1))(*[r.ptr for r in round_outputs])
        lens = (C.c_uint64 * max(n, 1))(*[r.len() for r in round_outputs])
        self._l._check(self._lib.b200_pairwise_product_reduce(self._ctx, input.ptr, input.len(), ptrs, lens, n))

    # -- fused round evaluation of v3::calculate_round_evals (what the traced accumulate_kernels
    #    program of core/src/protocols/sumcheck/v3/bivariate_product.rs:303-408 lowers to)
    def bivariate_round_evals(self, multilins: Sequence[DevSlice], n_vars: int, pairs, batch_coeff: int):
        m = len(multilins)
        for ml in multilins:
            if ml.len() != 1 << n_vars:
                raise InputValidation("every multilinear must have 2^n_vars elements")
        ptrs = (C.c_void_p * max(m, 1))(*[x.ptr for x in multilins])
        ia = (C.c_uint32 * max(len(pairs), 1))(*[p[0] for p in pairs])
        ib = (C.c_uint32 * max(len(pairs), 1))(*[p[1] for p in pairs])
        s1, s2 = C.c_uint32(), C.c_uint32()
        self._l._check(self._lib.b200_bivariate_round_evals(self._ctx, ptrs, m, n_vars, ia, ib, len(pairs), _u64x2(batch_coeff), C.byref(s1), C.byref(s2)))
        return OpValue(s1.value), OpValue(s2.value)


class B200Layer:
    """ComputeLayer<BinaryField128b> (layer.rs:22-88) on one B200."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        ctx = C.c_void_p()
        rc = self._lib.b200_ctx_create(device, C.byref(ctx))
        if rc != 0:
            raise DeviceError(f"b200_ctx_create(device={device}) failed with status {rc}: no usable sm_100 GPU "
                              "(binius_b200 has no CPU fallback)")
        self._ctx = ctx
        self._owned = []
        self._scratch = {}
        self._scratch_used = {}
        # `execute` is a scope over the context's result slots (reset at its start, fetched at its end): executes issued
        # from several host threads on one layer (`&self` in the reference, layer.rs:90-99) take turns; every other call
        # is serialised by the context's own lock inside the library
        import threading

        self._exec_lock = threading.RLock()

    def _check(self, rc: int):
        if rc != 0:
            _raise(rc, self._lib.b200_last_error(self._ctx).decode())

    # -- device memory owned by the holder / tests
    def dev_alloc(self, n: int) -> DevSlice:
        p = C.c_void_p()
        self._check(self._lib.b200_dev_alloc(self._ctx, n, C.byref(p)))
        self._owned.append(p.value)
        return DevSlice(p.value, n, True)

    def dev_free(self, s: DevSlice):
        self._check(self._lib.b200_dev_free(self._ctx, s.ptr))
        self._owned.remove(s.ptr)

    def _scratch_alloc(self, n: int) -> DevSlice:
        free = self._scratch.setdefault(n, [])
        s = free.pop() if free else self.dev_alloc(n)
        self._scratch_used.setdefault(n, []).append(s)
        return s

    def _scratch_release(self):
        for n, used in self._scratch_used.items():
            self._scratch.setdefault(n, []).extend(used)
        self._scratch_used = {}

    # -- ComputeLayer
    def copy_h2d(self, src: np.ndarray, dst: DevSlice):
        src = np.ascontiguousarray(src, dtype=np.uint64)
        n = src.size // 2
        if n != dst.len():
            raise InputValidation("precondition: src and dst buffers must have the same length")
        self._check(self._lib.b200_copy_h2d(self._ctx, src.ctypes.data, dst.ptr, n))
        self._check(self._lib.b200_sync(self._ctx))  # src may be a temporary

    def copy_d2h(self, src: DevSlice, dst: np.ndarray):
        assert dst.dtype == np.uint64 and dst.flags["C_CONTIGUOUS"]
        if dst.size // 2 != src.len():
            raise InputValidation("precondition: src and dst buffers must have the same length")
        self._check(self._lib.b200_copy_d2h(self._ctx, src.ptr, dst.ctypes.data, src.len()))

    def copy_d2d(self, src: DevSlice, dst: DevSlice):
        if src.len() != dst.len():
            raise InputValidation("precondition: src and dst buffers must have the same length")
        self._check(self._lib.b200_copy_d2d(self._ctx, src.ptr, dst.ptr, src.len()))

    def compile_expr(self, expr: ArithCircuit) -> ExprEval:
        return ExprEval(self, expr)

    def execute(self, f: Callable[[B200Executor], List[OpValue]]) -> List[int]:
        with self._exec_lock:
            self._check(self._lib.b200_results_reset(self._ctx))
            ex = B200Executor(self)
            try:
                vals = list(f(ex) or [])
                n = len(vals)
                slots = (C.c_uint32 * max(n, 1))(*[v.slot for v in vals])
                out = (C.c_uint64 * max(2 * n, 2))()
                self._check(self._lib.b200_results_fetch(self._ctx, slots, n, out))
            finally:
                self._scratch_release()
            return [int(out[2 * i]) | (int(out[2 * i + 1]) << 64) for i in range(n)]

    def fill(self, slice: DevSlice, value: int):
        self._check(self._lib.b200_fill(self._ctx, slice.ptr, slice.len(), _u64x2(value)))

    def extrapolate_line_host(self, evals_0: np.ndarray, evals_1: np.ndarray, z: int):
        """fold-high on HOST arrays (in place on evals_0): the old-HAL-shaped call
        (`Backend::Vec<P>` is host memory, hal/src/backend.rs:19-31); pipelined H2D/kernel/D2H."""
        if evals_0.shape != evals_1.shape:
            raise InputValidation("evals_0 and evals_1 must be the same length")
        assert evals_0.dtype == np.uint64 and evals_0.flags["C_CONTIGUOUS"] and evals_1.flags["C_CONTIGUOUS"]
        self._check(self._lib.b200_extrapolate_line_host(self._ctx, evals_0.ctypes.data, evals_1.ctypes.data, evals_0.size // 2, _u64x2(z)))

    def host_alloc(self, n_elems: int) -> np.ndarray:
        """pinned host buffer of n_elems B128 elements as an (n,2) uint64 array (never freed before close)"""
        p = C.c_void_p()
        self._check(self._lib.b200_host_alloc(self._ctx, n_elems * 16, C.byref(p)))
        return np.ctypeslib.as_array((C.c_uint64 * (2 * n_elems)).from_address(p.value)).reshape(n_elems, 2)

    # -- helpers for tests / bench
    def sync(self):
        self._check(self._lib.b200_sync(self._ctx))

    def launch_count(self) -> int:
        return int(self._lib.b200_ctx_launch_count(self._ctx))

    def set_tuning(self, key: str, value: int):
        """kernel-selection switch for A/B measurements and tests (b200_ctx_set_tuning)"""
        self._check(self._lib.b200_ctx_set_tuning(self._ctx, key.encode(), int(value)))

    def to_device(self, host: np.ndarray) -> DevSlice:
        host = np.ascontiguousarray(host, dtype=np.uint64).reshape(-1, 2)
        d = self.dev_alloc(len(host))
        self.copy_h2d(host, d)
        return d

    def to_host(self, d: DevSlice) -> np.ndarray:
        out = np.empty((d.len(), 2), dtype=np.uint64)
        self.copy_d2h(d, out)
        return out

    def close(self):
        if getattr(self, "_ctx", None):
            for p in self._owned:
                self._lib.b200_dev_free(self._ctx, p)
            self._owned = []
            self._lib.b200_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class ComputeData:
    """layer.rs:747-776"""
    hal: B200Layer
    host_alloc: HostBumpAllocator
    dev_alloc: BumpAllocator


class B200LayerHolder:
    """ComputeHolder (layer.rs:732-745); cf. CpuLayerHolder::new(host_mem_size, dev_mem_size)
    (compute/src/cpu/layer.rs:684-694) and FastCpuLayerHolder (examples/keccak.rs:119-122)."""

    def __init__(self, host_mem_size: int, dev_mem_size: int, device: int = 0):
        self.layer = B200Layer(device)
        self.host_mem = np.zeros((host_mem_size, 2), dtype=np.uint64)
        self.dev_mem = self.layer.dev_alloc(dev_mem_size)
        self.layer.fill(self.dev_mem, 0)

    @classmethod
    def new(cls, host_mem_size: int, dev_mem_size: int, device: int = 0):
        return cls(host_mem_size, dev_mem_size, device)

    def to_data(self) -> ComputeData:
        return ComputeData(self.layer, HostBumpAllocator(self.host_mem), BumpAllocator(self.dev_mem))


def calculate_round_evals(hal: B200Layer, n_vars: int, batch_coeff: int, multilins: Sequence[DevSlice], compositions) -> List[int]:
    """v3::calculate_round_evals (core/src/protocols/sumcheck/v3/bivariate_product.rs:303-408), literally: the
    accumulate_kernels program the reference prover issues -- per composition (an index pair = IndexComposition<
    BivariateProduct>, expression Var(i0) * Var(i1) over ALL multilinears) a sum over the high halves, add(lo, hi) into
    the Local buffers, the same sums over the Locals.  Returns [y_1, y_inf]."""
    m = len(multilins)
    evaluators = [hal.compile_expr(ArithCircuit.var(i0) * ArithCircuit.var(i1)) for (i0, i1) in compositions]
    split_n_vars = n_vars - 1
    mappings = []
    for ml in multilins:
        lo, hi = ml.split_half()
        mappings += [KernelMemMap.Chunked(lo, 0), KernelMemMap.Chunked(hi, 0), KernelMemMap.Local(split_n_vars)]
    coeffs, pw = [], 1
    from .hostfield import mul as _fmul  # batch-coefficient powers (host scalar work)

    for _ in compositions:
        coeffs.append(pw)
        pw = _fmul(pw, batch_coeff)

    def kernel(kex, log_chunks, buffers):
        log_chunk_size = split_n_vars - log_chunks
        acc_1 = kex.decl_value(0)
        eval_1s = SlicesBatch([buffers[3 * i + 1].to_ref() for i in range(m)], 1 << log_chunk_size)
        for c, ev in zip(coeffs, evaluators):
            kex.sum_composition_evals(eval_1s, ev, c, acc_1)
        for i in range(m):
            kex.add(log_chunk_size, buffers[3 * i].to_ref(), buffers[3 * i + 1].to_ref(), buffers[3 * i + 2].data)
        acc_inf = kex.decl_value(0)
        eval_infs = SlicesBatch([buffers[3 * i + 2].to_ref() for i in range(m)], 1 << log_chunk_size)
        for c, ev in zip(coeffs, evaluators):
            kex.sum_composition_evals(eval_infs, ev, c, acc_inf)
        return [acc_1, acc_inf]

    return hal.execute(lambda ex: ex.accumulate_kernels(kernel, mappings))


def eq_ind_partial_eval(hal: B200Layer, dev_alloc: BumpAllocator, point: Sequence[int]) -> DevSlice:
    """compute/src/ops.rs:26-50"""
    out = dev_alloc.alloc(1 << len(point))
    hal.fill(out.slice(0, 1), 1)
    hal.execute(lambda ex: (ex.tensor_expand(0, point, out), [])[1])
    return out
