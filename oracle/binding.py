"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under binius_b200/ may import this module.

B128 vectors are numpy uint64 arrays of shape (n, 2): [:, 0] = low 64 bits, [:, 1] = high 64 bits
(the little-endian u128 layout of BinaryField128b, reference crates/field/src/binary_field.rs:115-133).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

INPUT_VALIDATION = 1


class OracleError(Exception):
    def __init__(self, code):
        super().__init__(f"oracle error code {code}")
        self.code = code


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
    return _LIB


# ------------------------------------------------------------------------------------------------
# conversions
def to_arr(vals) -> np.ndarray:
    """list of python ints (< 2^128) -> (n,2) uint64"""
    a = np.empty((len(vals), 2), dtype=np.uint64)
    for i, v in enumerate(vals):
        a[i, 0] = v & 0xFFFFFFFFFFFFFFFF
        a[i, 1] = v >> 64
    return a


def to_ints(arr) -> list:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 2)
    return [int(lo) | (int(hi) << 64) for lo, hi in arr]


def one(v: int) -> np.ndarray:
    return to_arr([v])


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.ndim == 2 and a.shape[1] == 2
    return a


def _check(rc):
    if rc != 0:
        raise OracleError(rc)


class ExprStep(C.Structure):
    _fields_ = [("op", C.c_uint32), ("l", C.c_uint32), ("r", C.c_uint64), ("c_lo", C.c_uint64), ("c_hi", C.c_uint64)]


def encode_expr(steps):
    """steps: list of ('add',l,r) | ('mul',l,r) | ('pow',l,e) | ('const',c) | ('var',i)"""
    arr = (ExprStep * max(len(steps), 1))()
    for i, st in enumerate(steps):
        k = st[0]
        if k == "add":
            arr[i] = ExprStep(0, st[1], st[2], 0, 0)
        elif k == "mul":
            arr[i] = ExprStep(1, st[1], st[2], 0, 0)
        elif k == "pow":
            arr[i] = ExprStep(2, st[1], st[2], 0, 0)
        elif k == "const":
            arr[i] = ExprStep(3, 0, 0, st[1] & 0xFFFFFFFFFFFFFFFF, st[1] >> 64)
        elif k == "var":
            arr[i] = ExprStep(4, st[1], 0, 0, 0)
        else:
            raise ValueError(k)
    return arr


def expr_n_vars(steps):
    return 1 + max([st[1] for st in steps if st[0] == "var"], default=-1)


# ------------------------------------------------------------------------------------------------
# scalar field ops (python ints)
def _scalar2(fn, a, b, k):
    out = one(0)
    getattr(lib(), fn)(_p(one(a)), _p(one(b)), C.c_uint32(k), _p(out))
    return to_ints(out)[0]


def _scalar1(fn, a, k):
    out = one(0)
    getattr(lib(), fn)(_p(one(a)), C.c_uint32(k), _p(out))
    return to_ints(out)[0]


def mul(a, b, k=7):
    return _scalar2("orc_mul", a, b, k)


def mul_slow(a, b, k=7):
    return _scalar2("orc_mul_slow", a, b, k)


def mul_subfield(a, s, k):
    return _scalar2("orc_mul_subfield", a, s, k)


def square(a, k=7):
    return _scalar1("orc_square", a, k)


def invert(a, k=7):
    return _scalar1("orc_invert", a, k)


def mul_alpha(a, k=7):
    return _scalar1("orc_mul_alpha", a, k)


def pow_(a, e, k=7):
    r = 1
    while e:
        if e & 1:
            r = mul(r, a, k)
        a = mul(a, a, k)
        e >>= 1
    return r


def mul_vec(a, b):
    a, b = _c(a), _c(b)
    out = np.empty_like(a)
    lib().orc_mul_vec(_p(a), _p(b), _p(out), C.c_uint64(len(a)))
    return out


# ------------------------------------------------------------------------------------------------
# ComputeLayer ops
def extrapolate_line(e0, e1, z: int):
    e0 = _c(e0).copy()
    e1 = _c(e1)
    if len(e0) != len(e1):
        raise OracleError(INPUT_VALIDATION)
    _check(lib().orc_extrapolate_line(_p(e0), _p(e1), C.c_uint64(len(e0)), _p(one(z))))
    return e0


def tensor_expand(data, log_n: int, coords):
    data = _c(data).copy()
    cs = to_arr(list(coords)) if len(coords) else np.zeros((1, 2), np.uint64)
    _check(lib().orc_tensor_expand(_p(data), C.c_uint64(len(data)), C.c_uint32(log_n), _p(cs), C.c_uint32(len(coords))))
    return data


def inner_product(a, lvl: int, b) -> int:
    a, b = _c(a), _c(b)
    out = one(0)
    _check(lib().orc_inner_product(_p(a), C.c_uint64(len(a)), C.c_uint32(lvl), _p(b), C.c_uint64(len(b)), _p(out)))
    return to_ints(out)[0]


def _fold(fn, mat, lvl, vec, n_out):
    mat, vec = _c(mat), _c(vec)
    out = np.zeros((n_out, 2), np.uint64)
    _check(getattr(lib(), fn)(_p(mat), C.c_uint64(len(mat)), C.c_uint32(lvl), _p(vec), C.c_uint64(len(vec)), _p(out), C.c_uint64(n_out)))
    return out


def fold_left(mat, lvl, vec, n_out):
    return _fold("orc_fold_left", mat, lvl, vec, n_out)


def fold_right(mat, lvl, vec, n_out):
    return _fold("orc_fold_right", mat, lvl, vec, n_out)


def _ptr_array(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def compute_composite(inputs, steps, n_out=None):
    inputs = [_c(a) for a in inputs]
    row_len = len(inputs[0]) if inputs else 0
    n_out = row_len if n_out is None else n_out
    out = np.zeros((n_out, 2), np.uint64)
    enc = encode_expr(steps)
    _check(lib().orc_compute_composite(_ptr_array(inputs), C.c_uint32(len(inputs)), C.c_uint64(row_len), _p(out),
                                       C.c_uint64(n_out), enc, C.c_uint32(len(steps)), C.c_uint32(expr_n_vars(steps))))
    return out


def sum_composition_evals(inputs, steps, batch_coeff: int, acc: int) -> int:
    inputs = [_c(a) for a in inputs]
    accv = one(acc)
    enc = encode_expr(steps)
    _check(lib().orc_sum_composition_evals(_ptr_array(inputs), C.c_uint32(len(inputs)), C.c_uint64(len(inputs[0])), enc,
                                           C.c_uint32(len(steps)), _p(one(batch_coeff)), _p(accv)))
    return to_ints(accv)[0]


def pairwise_product_reduce(inp, out_lens=None):
    inp = _c(inp)
    n = len(inp)
    if out_lens is None:
        out_lens = [n >> (r + 1) for r in range(max(n.bit_length() - 1, 0))]
    outs = [np.zeros((max(l, 1), 2), np.uint64)[:l] for l in out_lens]
    bufs = [np.ascontiguousarray(o) if len(o) else np.zeros((1, 2), np.uint64) for o in outs]
    lens = (C.c_uint64 * max(len(out_lens), 1))(*out_lens)
    _check(lib().orc_pairwise_product_reduce(_p(inp), C.c_uint64(n), _ptr_array(bufs) if bufs else None, lens, C.c_uint32(len(out_lens))))
    return [b[:l] for b, l in zip(bufs, out_lens)]


def bivariate_round_evals(multilins, n_vars, pairs, batch_coeff: int):
    multilins = [_c(m) for m in multilins]
    ia = (C.c_uint32 * len(pairs))(*[p[0] for p in pairs])
    ib = (C.c_uint32 * len(pairs))(*[p[1] for p in pairs])
    out = np.zeros((2, 2), np.uint64)
    _check(lib().orc_bivariate_round_evals(_ptr_array(multilins), C.c_uint32(len(multilins)), C.c_uint32(n_vars), ia, ib,
                                           C.c_uint32(len(pairs)), _p(one(batch_coeff)), _p(out)))
    return to_ints(out)


# ------------------------------------------------------------------------------------------------
# timed CPU arm (oracle/cpu_baseline.c): C restatement of the reference's GFNI/AVX-512 fold path
def cpu_fold(e0, e1, z: int, n_threads: int = 1, use_gfni: bool = True):
    a = _c(e0).copy()
    used = lib().cpu_fold(_p(a), _p(_c(e1)), C.c_uint64(len(a)), _p(one(z)), C.c_int(n_threads), C.c_int(int(use_gfni)))
    return a, bool(used)


def cpu_fold_parallel(log_coeffs: int, budget_s: float = 10.0, n_threads: int = 0):
    """Times fold-high over a bounded sample on this host; returns the bench.py cpu_baseline object."""
    import os

    n_threads = n_threads or (os.cpu_count() or 1)
    log_s = min(log_coeffs, 24)
    fn = lib().cpu_fold_bench
    fn.restype = C.c_double
    used = C.c_int()
    t1 = fn(C.c_uint32(log_s), C.c_int(1), C.c_int(n_threads), C.c_int(1), C.byref(used))
    reps = max(1, min(200, int(budget_s / max(t1, 1e-4))))
    t = fn(C.c_uint32(log_s), C.c_int(reps), C.c_int(n_threads), C.c_int(1), C.byref(used))
    how = "AVX-512+GFNI" if used.value else "scalar (host CPU lacks AVX-512/GFNI)"
    return {"value": reps * (1 << log_s) / t, "unit": "coeffs/s", "cores": n_threads, "kind": "port",
            "sample": f"2^{log_s} coefficients x {reps} reps, {n_threads} threads, {how}; C restatement of the reference "
                      f"GFNI fold path (fold_left_lerp_inplace), not the Rust binary"}


def cpu_fold_chain_parallel(log_coeffs: int, budget_s: float = 6.0, n_threads: int = 0):
    """Times the complete fold chain (log_coeffs rounds of fold-high, 2^(log_coeffs+1) - 2 coefficients in
    total) of one multilinear on this host; returns {"value": coeffs/s, ...}."""
    import os

    n_threads = n_threads or (os.cpu_count() or 1)
    fn = lib().cpu_fold_chain_bench
    fn.restype = C.c_double
    t1 = fn(C.c_uint32(log_coeffs), C.c_int(1), C.c_int(n_threads), C.c_int(1))
    reps = max(1, min(50, int(budget_s / max(t1, 1e-4))))
    t = fn(C.c_uint32(log_coeffs), C.c_int(reps), C.c_int(n_threads), C.c_int(1))
    total = (2 << log_coeffs) - 2
    return {"value": reps * total / t, "unit": "coeffs/s", "ms_per_chain": t / reps * 1e3, "cores": n_threads, "kind": "port",
            "sample": f"{reps} x full fold chain of a 2^{log_coeffs}-coefficient multilinear, {n_threads} threads"}


def cpu_ntt_forward(data, log_x: int, log_y: int, skip: int = 0, d: int = 24, n_threads: int = 1, use_gfni: bool = True):
    """Threaded GFNI forward NTT over B32 (oracle/cpu_baseline.c, port of crates/ntt/src/multithreaded.rs:100-228),
    shape (log_x >= 4, log_y, 0), coset 0; returns the transformed copy."""
    a = np.ascontiguousarray(data, dtype=np.uint32).copy()
    assert a.shape[0] == 1 << (log_x + log_y)
    ntt = NTT(5, d)
    rc = lib().cpu_ntt_forward(_p(a), C.c_uint32(log_x), C.c_uint32(log_y), C.c_uint32(skip), _p(ntt.s), C.c_uint32(d),
                               C.c_int(n_threads), C.c_int(int(use_gfni)))
    _check(rc)
    return a


def cpu_ntt_parallel(log_x: int, log_y: int, skip: int, budget_s: float = 6.0, n_threads: int = 0, d: int = 0):
    """Times the forward NTT of 2^(log_x+log_y) B32 coefficients on this host; returns a cpu_baseline object."""
    import os

    n_threads = n_threads or (os.cpu_count() or 1)
    d = d or max(log_y, 1)
    ntt = NTT(5, d)
    fn = lib().cpu_ntt_bench
    fn.restype = C.c_double
    args = (C.c_uint32(log_x), C.c_uint32(log_y), C.c_uint32(skip))
    t1 = fn(*args, C.c_int(1), C.c_int(n_threads), _p(ntt.s), C.c_uint32(d), C.c_int(1))
    reps = max(1, min(100, int(budget_s / max(t1, 1e-4))))
    t = fn(*args, C.c_int(reps), C.c_int(n_threads), _p(ntt.s), C.c_uint32(d), C.c_int(1))
    n = 1 << (log_x + log_y)
    how = "AVX-512+GFNI" if lib().cpu_has_gfni512() else "scalar (host CPU lacks AVX-512/GFNI)"
    return {"value": reps * n / t, "unit": "coeffs/s", "ms_per_transform": t / reps * 1e3, "cores": n_threads, "kind": "port",
            "sample": f"{reps} x forward NTT (log_x {log_x}, log_y {log_y}, skip {skip}) of 2^{log_x + log_y} B32 coefficients, "
                      f"{n_threads} threads, {how}; C restatement of crates/ntt/src/multithreaded.rs, not the Rust binary"}


def cpu_bivariate_round_evals(multilins, n_vars: int, pairs, batch_coeff: int, n_threads: int = 1, use_gfni: bool = True):
    """[y_1, y_inf] by the threaded GFNI CPU arm (oracle/cpu_baseline.c)"""
    ms = [_c(x) for x in multilins]
    ptrs = (C.c_void_p * len(ms))(*[x.ctypes.data for x in ms])
    ia = (C.c_uint32 * max(len(pairs), 1))(*[p[0] for p in pairs])
    ib = (C.c_uint32 * max(len(pairs), 1))(*[p[1] for p in pairs])
    out = np.zeros((2, 2), np.uint64)
    lib().cpu_bivariate_round_evals(ptrs, C.c_uint32(n_vars), ia, ib, C.c_uint32(len(pairs)), _p(one(batch_coeff)), C.c_int(n_threads),
                                    C.c_int(int(use_gfni)), _p(out))
    return to_ints(out)


def cpu_bivariate_sumcheck_parallel(m: int, n_vars: int, n_comp: int, budget_s: float = 6.0, n_threads: int = 0):
    """Times the data plane of a whole bivariate-product sumcheck (n_vars rounds of round evaluations + fold of every
    multilinear) on this host; returns a cpu_baseline object in rounds/s."""
    import os

    n_threads = n_threads or (os.cpu_count() or 1)
    fn = lib().cpu_bivariate_sumcheck_bench
    fn.restype = C.c_double
    chk = np.zeros((1, 2), np.uint64)
    args = (C.c_uint32(m), C.c_uint32(n_vars), C.c_uint32(n_comp))
    t1 = fn(*args, C.c_int(1), C.c_int(n_threads), C.c_int(1), _p(chk))
    reps = max(1, min(20, int(budget_s / max(t1, 1e-4))))
    t = fn(*args, C.c_int(reps), C.c_int(n_threads), C.c_int(1), _p(chk))
    return {"value": reps * n_vars / t, "unit": "rounds/s", "ms_per_sumcheck": t / reps * 1e3, "cores": n_threads, "kind": "port",
            "sample": f"{reps} x sumcheck over {m} multilinears of 2^{n_vars} B128 elements, {n_comp} index pairs, {n_threads} threads; "
                      "C restatement of v3::BivariateSumcheckProver over FastCpuLayer (GFNI multiply), not the Rust binary"}


def cpu_chi_round_evals(cols, n_out: int, n_b: int, n_vars: int, eq, with_eval_1: bool, n_threads: int = 1, use_gfni: bool = True):
    """[constraint][at 1, at infinity] of the keccak chi constraints out_c - (b_c + (b_{c+1} - 1) * b_{c+2}) by the threaded
    GFNI CPU arm; cols = n_out `out` columns followed by n_b `b` columns of 2^n_vars B128 elements"""
    cs = [_c(x) for x in cols]
    ptrs = (C.c_void_p * len(cs))(*[x.ctypes.data for x in cs])
    out = np.zeros((2 * n_out, 2), np.uint64)
    lib().cpu_chi_round_evals(ptrs, C.c_uint32(n_out), C.c_uint32(n_b), C.c_uint32(n_vars), _p(_c(eq)), C.c_int(int(with_eval_1)),
                              C.c_int(n_threads), C.c_int(int(use_gfni)), _p(out))
    v = to_ints(out)
    return [[v[2 * c], v[2 * c + 1]] for c in range(n_out)]


def cpu_chi_zerocheck_parallel(n_out: int, n_b: int, n_vars: int, budget_s: float = 6.0, n_threads: int = 0):
    """Times all rounds of the chi-constraint zerocheck (round values + folds + eq halving) on this host."""
    import os

    n_threads = n_threads or (os.cpu_count() or 1)
    fn = lib().cpu_chi_zerocheck_bench
    fn.restype = C.c_double
    chk = np.zeros((1, 2), np.uint64)
    args = (C.c_uint32(n_out), C.c_uint32(n_b), C.c_uint32(n_vars))
    t1 = fn(*args, C.c_int(1), C.c_int(n_threads), C.c_int(1), _p(chk))
    reps = max(1, min(10, int(budget_s / max(t1, 1e-4))))
    t = fn(*args, C.c_int(reps), C.c_int(n_threads), C.c_int(1), _p(chk)) if reps > 1 else t1
    return {"value": reps * n_vars / t, "unit": "rounds/s", "ms_per_sumcheck": t / reps * 1e3, "cores": n_threads, "kind": "port",
            "sample": f"{reps} x zerocheck rounds over {n_out + n_b} multilinears of 2^{n_vars} B128 elements, {n_out} chi constraints, "
                      f"{n_threads} threads; C restatement of the eq-ind evaluator loop (GFNI multiply), not the Rust binary"}


def cpu_tensor_expand(data, log_n: int, coords, n_threads: int = 1, use_gfni: bool = True):
    """the threaded GFNI CPU arm of tensor_expand: data[: 2^log_n] filled, returns the 2^(log_n + k) expansion"""
    k = len(coords)
    buf = np.zeros(((1 << (log_n + k)), 2), np.uint64)
    buf[: 1 << log_n] = _c(data)[: 1 << log_n]
    cs = to_arr(list(coords)) if not isinstance(coords, np.ndarray) else _c(coords)
    lib().cpu_tensor_expand(_p(buf), C.c_uint32(log_n), _p(cs), C.c_uint32(k), C.c_int(n_threads), C.c_int(int(use_gfni)))
    return buf


def cpu_tensor_expand_parallel(k: int, budget_s: float = 3.0, n_threads: int = 0):
    """Times the expansion of [1] by k coordinates (tensor_product_full_query) on this host."""
    import os

    n_threads = n_threads or (os.cpu_count() or 1)
    fn = lib().cpu_tensor_expand_bench
    fn.restype = C.c_double
    chk = np.zeros((1, 2), np.uint64)
    t1 = fn(C.c_uint32(k), C.c_int(1), C.c_int(n_threads), C.c_int(1), _p(chk))
    reps = max(1, min(200, int(budget_s / max(t1, 1e-4))))
    t = fn(C.c_uint32(k), C.c_int(reps), C.c_int(n_threads), C.c_int(1), _p(chk)) if reps > 1 else t1
    return {"value": reps * (1 << k) / t, "unit": "elems/s", "ms": t / reps * 1e3, "cores": n_threads, "kind": "port",
            "sample": f"{reps} x expansion of [1] by {k} coordinates (2^{k} B128 elements), {n_threads} threads, AVX-512+GFNI; "
                      "C restatement of tensor_prod_eq_ind (crates/math/src/tensor_prod_eq_ind.rs:35-77), not the Rust binary"}


def cpu_u32add_round_evals(cols, n_vars: int, eq, with_eval_1: bool, n_threads: int = 1, use_gfni: bool = True):
    """[C1 at 1, C1 at infinity, C2 at 1] of the u32_add gadget compositions (BASELINE config #3) by the threaded GFNI
    CPU arm; cols = x, y, cin, cout, z as 2^n_vars B128 elements each"""
    cs = [_c(x) for x in cols]
    ptrs = (C.c_void_p * len(cs))(*[x.ctypes.data for x in cs])
    out = np.zeros((3, 2), np.uint64)
    lib().cpu_u32add_round_evals(ptrs, C.c_uint32(n_vars), _p(_c(eq)), C.c_int(int(with_eval_1)), C.c_int(n_threads), C.c_int(int(use_gfni)), _p(out))
    return to_ints(out)


def cpu_u32add_zerocheck_parallel(n_vars: int, budget_s: float = 5.0, n_threads: int = 0):
    """Times all rounds of the u32_add zerocheck (round values + folds + eq halving) on this host."""
    import os

    n_threads = n_threads or (os.cpu_count() or 1)
    fn = lib().cpu_u32add_zerocheck_bench
    fn.restype = C.c_double
    chk = np.zeros((1, 2), np.uint64)
    t1 = fn(C.c_uint32(n_vars), C.c_int(1), C.c_int(n_threads), C.c_int(1), _p(chk))
    reps = max(1, min(200, int(budget_s / max(t1, 1e-4))))
    t = fn(C.c_uint32(n_vars), C.c_int(reps), C.c_int(n_threads), C.c_int(1), _p(chk)) if reps > 1 else t1
    return {"value": reps * n_vars / t, "unit": "rounds/s", "ms_per_sumcheck": t / reps * 1e3, "cores": n_threads, "kind": "port",
            "sample": f"{reps} x all {n_vars} zerocheck rounds over the 5 u32_add multilinears of 2^{n_vars} B128 elements, {n_threads} threads; "
                      "C restatement of the eq-ind evaluator loop (GFNI multiply), not the Rust binary"}


# ------------------------------------------------------------------------------------------------
# old HAL (ComputationBackend) restatements, oracle/hal.c
def fold_left_lerp_inplace(evals, prefix: int, suffix: int, log_n: int, z: int):
    """returns the folded (truncated) vector"""
    buf = _c(evals).copy()
    lib().orc_fold_left_lerp_inplace.restype = C.c_uint64
    n = lib().orc_fold_left_lerp_inplace(_p(buf), C.c_uint64(prefix), _p(one(suffix)), C.c_uint32(log_n), _p(one(z)))
    return buf[:n]


def eq_ind_round_evals(mls, lens, suffixes, n_vars, eq_ind, comps, leads, codes, points):
    """comps / leads: lists of step lists; returns [[value per code] per composition]"""
    mls = [_c(m) if len(m) else np.zeros((1, 2), np.uint64) for m in mls]
    m = len(mls)
    enc_c = [encode_expr(s) for s in comps]
    enc_l = [encode_expr(s) for s in leads]
    pc = (C.c_void_p * len(comps))(*[C.addressof(e) for e in enc_c])
    pl = (C.c_void_p * len(comps))(*[C.addressof(e) for e in enc_l])
    nc = (C.c_uint32 * len(comps))(*[len(s) for s in comps])
    nl = (C.c_uint32 * len(comps))(*[len(s) for s in leads])
    out = np.zeros((max(len(comps) * len(codes), 1), 2), np.uint64)
    lib().orc_eq_ind_round_evals(_ptr_array(mls), (C.c_uint64 * max(m, 1))(*lens), _p(to_arr(list(suffixes)) if m else np.zeros((1, 2), np.uint64)),
                                 C.c_uint32(m), C.c_uint32(n_vars), _p(_c(eq_ind)), pc, nc, pl, nl, C.c_uint32(len(comps)),
                                 (C.c_uint32 * max(len(codes), 1))(*codes), _p(to_arr(list(points)) if len(points) else np.zeros((1, 2), np.uint64)),
                                 C.c_uint32(len(codes)), _p(out))
    vals = to_ints(out)
    return [vals[c * len(codes):(c + 1) * len(codes)] for c in range(len(comps))]


def sumcheck_round_evals(order: int, mls, lens, suffixes, n_vars, eq_ind, comps, leads, codes, points):
    """order 0 = LowToHigh, 1 = HighToLow; eq_ind None = regular (unweighted) evaluator."""
    mls = [_c(m) if len(m) else np.zeros((1, 2), np.uint64) for m in mls]
    m = len(mls)
    enc_c = [encode_expr(s) for s in comps]
    enc_l = [encode_expr(s) for s in leads]
    pc = (C.c_void_p * len(comps))(*[C.addressof(e) for e in enc_c])
    pl = (C.c_void_p * len(comps))(*[C.addressof(e) for e in enc_l])
    nc = (C.c_uint32 * len(comps))(*[len(s) for s in comps])
    nl = (C.c_uint32 * len(comps))(*[len(s) for s in leads])
    out = np.zeros((max(len(comps) * len(codes), 1), 2), np.uint64)
    eq = _p(_c(eq_ind)) if eq_ind is not None else C.c_void_p(None)
    lib().orc_sumcheck_round_evals(C.c_uint32(order), _ptr_array(mls), (C.c_uint64 * max(m, 1))(*lens),
                                   _p(to_arr(list(suffixes)) if m else np.zeros((1, 2), np.uint64)),
                                   C.c_uint32(m), C.c_uint32(n_vars), eq, pc, nc, pl, nl, C.c_uint32(len(comps)),
                                   (C.c_uint32 * max(len(codes), 1))(*codes), _p(to_arr(list(points)) if len(points) else np.zeros((1, 2), np.uint64)),
                                   C.c_uint32(len(codes)), _p(out))
    vals = to_ints(out)
    return [vals[c * len(codes):(c + 1) * len(codes)] for c in range(len(comps))]


def fold_right_lerp(evals, suffix: int, z: int):
    """fold_right_lerp over the stored prefix `evals`; returns the ceil(len/2) folded elements"""
    buf = _c(evals) if len(evals) else np.zeros((1, 2), np.uint64)
    out = np.zeros(((len(evals) + 1) // 2 + 1, 2), np.uint64)
    lib().orc_fold_right_lerp.restype = C.c_uint64
    n = lib().orc_fold_right_lerp(_p(buf), C.c_uint64(len(evals)), _p(one(suffix)), _p(one(z)), _p(out))
    return out[:n]


def fold_partial_eq_ind_low_to_high(e):
    buf = _c(e)
    out = np.zeros((max(len(buf) // 2, 1), 2), np.uint64)
    lib().orc_fold_partial_eq_ind_low_to_high(_p(buf), C.c_uint64(len(buf)), _p(out))
    return out[: len(buf) // 2]


def fold_partial_eq_ind(e):
    buf = _c(e).copy()
    lib().orc_fold_partial_eq_ind(_p(buf), C.c_uint64(len(buf)))
    return buf[: len(buf) // 2]


# ------------------------------------------------------------------------------------------------
# zerocheck univariate-skip round (oracle/univariate.c): the specification of SURVEY.md 8f rank 1
def lagrange_evals(k: int, x: int):
    out = np.zeros((1 << k, 2), np.uint64)
    lib().orc_lagrange_evals(C.c_uint32(k), _p(one(x)), _p(out))
    return to_ints(out)


def zerocheck_univariate_evals(mls, levels, n_vars: int, skip: int, eq_ind, comps, max_domain_size: int):
    """mls: packed sub-field multilinears (arrays of B128 words), levels: their tower levels; returns
    [[R_c(x_i) for i < max_domain_size - 2^skip] for each composition]."""
    mls = [_c(x) for x in mls]
    m = len(mls)
    enc = [encode_expr(s) for s in comps]
    pc = (C.c_void_p * len(comps))(*[C.addressof(e) for e in enc])
    nc = (C.c_uint32 * len(comps))(*[len(s) for s in comps])
    n_points = max_domain_size - (1 << skip)
    out = np.zeros((max(len(comps) * n_points, 1), 2), np.uint64)
    rc = lib().orc_zerocheck_univariate_evals(_ptr_array(mls), (C.c_uint32 * m)(*levels), C.c_uint32(m), C.c_uint32(n_vars), C.c_uint32(skip),
                                              _p(_c(eq_ind)), pc, nc, C.c_uint32(len(comps)), C.c_uint32(max_domain_size), _p(out))
    _check(rc)
    vals = to_ints(out)
    return [vals[c * n_points:(c + 1) * n_points] for c in range(len(comps))]


def extrapolate_round_evals(staggered, skip: int, degree: int, max_domain_size: int):
    """univariate.rs:565-640: `staggered` = the (degree-1)*2^skip evaluations after the skipped domain (a longer
    list is truncated to them); returns the max_domain_size - 2^skip round evals the reference outputs."""
    n_points = max_domain_size - (1 << skip)
    n_in = max(degree - 1, 0) << skip
    vals = np.zeros((max(n_points, 1), 2), np.uint64)
    if n_in:
        vals[:n_in] = to_arr(list(staggered[:n_in]))
    lib().orc_extrapolate_round_evals(C.c_uint32(skip), C.c_uint32(degree), C.c_uint32(max_domain_size), _p(vals))
    return to_ints(vals)[:n_points]


def zerocheck_univariate_evals_reference(mls, levels, n_vars, skip, eq_ind, comps, degrees, max_domain_size):
    """What the reference's zerocheck_univariate_evals returns (univariate.rs:235-500): each composition
    evaluated at its (deg-1)*2^skip points, then extended to max_domain_size assuming zeros on the skipped domain."""
    full = zerocheck_univariate_evals(mls, levels, n_vars, skip, eq_ind, comps, max_domain_size)
    return [extrapolate_round_evals(v, skip, d, max_domain_size) for v, d in zip(full, degrees)]


def expand_monomials(steps, max_degree=2):
    """ArithCircuit steps -> {sorted tuple of variable indices: coefficient} (polynomials of degree <= max_degree)"""
    polys = []
    for st in steps:
        k = st[0]
        if k == "const":
            p = {(): st[1]} if st[1] else {}
        elif k == "var":
            p = {(st[1],): 1}
        elif k == "add":
            p = dict(polys[st[1]])
            for mono, c in polys[st[2]].items():
                v = p.get(mono, 0) ^ c
                if v:
                    p[mono] = v
                else:
                    p.pop(mono, None)
        elif k in ("mul", "pow"):
            factors = [polys[st[1]], polys[st[2]]] if k == "mul" else [polys[st[1]]] * st[2]
            p = {(): 1}
            for f in factors:
                nxt = {}
                for m1, c1 in p.items():
                    for m2, c2 in f.items():
                        mono = tuple(sorted(m1 + m2))
                        if len(mono) > max_degree:
                            raise ValueError("degree above max_degree")
                        v = nxt.get(mono, 0) ^ mul(c1, c2)
                        if v:
                            nxt[mono] = v
                        else:
                            nxt.pop(mono, None)
                p = nxt
        else:
            raise ValueError(k)
        polys.append(p)
    return polys[-1] if polys else {}


def cpu_univariate_b1(cols, n_vars: int, skip: int, eq_ind, comps, n_pts: int, n_threads: int = 0):
    """Threaded CPU arm of the univariate-skip round (oracle/cpu_univariate.c): B1 columns, compositions of degree <= 2
    with B8 constants, every composition evaluated at the n_pts points after the skipped domain.
    Returns (round evals [composition][point], seconds inside the C call)."""
    import os
    import time

    cols = [np.ascontiguousarray(c).view(np.uint8).reshape(-1) for c in cols]
    ma, mb, mc, first, cnt = [], [], [], [], []
    for steps in comps:
        poly = expand_monomials(steps)
        first.append(len(ma))
        for mono, c in poly.items():
            if c >> 8:
                raise ValueError("constant outside B8")
            ma.append(mono[0] if len(mono) > 0 else 0xFFFF)
            mb.append(mono[1] if len(mono) > 1 else 0xFFFF)
            mc.append(c)
        cnt.append(len(ma) - first[-1])
    nm, nc = max(len(ma), 1), len(comps)
    out = np.zeros((max(nc * n_pts, 1), 2), np.uint64)
    ptrs = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    n_threads = n_threads or len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    rc = lib().orc_cpu_univariate_b1(ptrs, C.c_uint32(len(cols)), C.c_uint32(n_vars), C.c_uint32(skip), _p(_c(eq_ind)),
                                     (C.c_uint16 * nm)(*ma), (C.c_uint16 * nm)(*mb), (C.c_uint8 * nm)(*mc),
                                     (C.c_uint32 * max(nc, 1))(*first), (C.c_uint32 * max(nc, 1))(*cnt), C.c_uint32(nc),
                                     C.c_uint32(n_pts), C.c_uint32(n_threads), _p(out))
    dt = time.perf_counter() - t0
    _check(rc)
    vals = to_ints(out)
    return [vals[c * n_pts:(c + 1) * n_pts] for c in range(nc)], dt


# ------------------------------------------------------------------------------------------------
# additive NTT
_NP_DT = {3: np.uint8, 4: np.uint16, 5: np.uint32, 6: np.uint64}


class NTT:
    """AdditiveNTT over T_kt with the standard subspace <1,2,4,...> of dimension d
    (reference: SingleThreadedNTT::new(log_domain_size), ntt/src/single_threaded.rs:27-45)."""

    def __init__(self, kt: int, d: int):
        self.kt, self.d = kt, d
        self.s = np.zeros((max(d * max(d - 1, 1), 1), 2), np.uint64)
        _check(lib().orc_ntt_s_evals(C.c_uint32(kt), C.c_uint32(d), _p(self.s)))

    def s_evals(self):
        """list of rows (python ints)"""
        W = self.d - 1
        flat = to_ints(self.s)
        return [flat[r * W: r * W + (self.d - 1 - r)] for r in range(self.d)]

    def get_subspace_eval(self, i: int, j: int) -> int:
        out = one(0)
        _check(lib().orc_ntt_get_subspace_eval(_p(self.s), C.c_uint32(self.d), C.c_uint32(i), C.c_uint64(j), _p(out)))
        return to_ints(out)[0]

    def _tr(self, inverse, data, kd, log_x, log_y, log_z, coset, coset_bits, skip_rounds):
        data = np.ascontiguousarray(data).copy()
        n_elems = data.shape[0]
        _check(lib().orc_ntt_transform(C.c_int(inverse), _p(self.s), C.c_uint32(self.kt), C.c_uint32(self.d), _p(data),
                                       C.c_uint32(kd), C.c_uint64(n_elems), C.c_uint32(log_x), C.c_uint32(log_y),
                                       C.c_uint32(log_z), C.c_uint64(coset), C.c_uint32(coset_bits), C.c_uint32(skip_rounds)))
        return data

    def forward(self, data, kd, log_x=0, log_y=None, log_z=0, coset=0, coset_bits=0, skip_rounds=0):
        if log_y is None:
            log_y = (data.shape[0]).bit_length() - 1 - log_x - log_z
        return self._tr(0, data, kd, log_x, log_y, log_z, coset, coset_bits, skip_rounds)

    def inverse(self, data, kd, log_x=0, log_y=None, log_z=0, coset=0, coset_bits=0, skip_rounds=0):
        if log_y is None:
            log_y = (data.shape[0]).bit_length() - 1 - log_x - log_z
        return self._tr(1, data, kd, log_x, log_y, log_z, coset, coset_bits, skip_rounds)

    def fri_fold(self, log_len, log_batch, challenges, data_in, n_out):
        data_in = _c(data_in)
        ch = to_arr(list(challenges)) if len(challenges) else np.zeros((1, 2), np.uint64)
        out = np.zeros((max(n_out, 1), 2), np.uint64)
        _check(lib().orc_fri_fold(_p(self.s), C.c_uint32(self.kt), C.c_uint32(self.d), C.c_uint32(log_len),
                                  C.c_uint32(log_batch), _p(ch), C.c_uint32(len(challenges)), _p(data_in),
                                  C.c_uint64(len(data_in)), _p(out), C.c_uint64(n_out)))
        return out[:n_out]


# ------------------------------------------------------------------------------------------------
# BinaryField128bPolyval (oracle/polyval.c)
def polyval_mul(a: int, b: int) -> int:
    out = one(0)
    lib().orc_polyval_mul(_p(one(a)), _p(one(b)), _p(out))
    return to_ints(out)[0]


def polyval_mul_vec(a, b):
    a, b = _c(a), _c(b)
    out = np.zeros_like(a)
    lib().orc_polyval_mul_vec(_p(a), _p(b), _p(out), C.c_uint64(len(a)))
    return out


def linear_map(images, x):
    """FieldLinearTransformation::transform with the 128 basis images (python ints)"""
    x = _c(x)
    out = np.zeros_like(x)
    lib().orc_linear_map(_p(to_arr(list(images))), _p(x), _p(out), C.c_uint64(len(x)))
    return out


def polyval_gpa_layers(inp, n_vars: int):
    """GrandProductWitness::new over POLYVAL elements: list of layers (2^n_vars, 2^(n_vars-1), ..., 1 elements)"""
    inp = _c(inp)
    assert len(inp) == 1 << n_vars
    out = np.zeros(((2 << n_vars) - 1, 2), np.uint64)
    lib().orc_polyval_gpa_layers(_p(inp), C.c_uint32(n_vars), _p(out))
    layers, off = [], 0
    for k in range(n_vars + 1):
        ln = 1 << (n_vars - k)
        layers.append(out[off:off + ln])
        off += ln
    return layers


def polyval_gpa_round_evals(a, b, eq, n_vars: int):
    out = np.zeros((2, 2), np.uint64)
    lib().orc_polyval_gpa_round_evals(_p(_c(a)), _p(_c(b)), _p(_c(eq)), C.c_uint32(n_vars), _p(out))
    return to_ints(out)


# ------------------------------------------------------------------------------------------------
# Groestl-256 / binary Merkle tree (oracle/groestl.c)
def groestl256(msg: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    buf = (C.c_uint8 * max(len(msg), 1)).from_buffer_copy(msg if msg else b"\0")
    lib().orc_groestl256(buf, C.c_uint64(len(msg)), out)
    return bytes(out)


def groestl256_compress_pair(left: bytes, right: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().orc_groestl256_compress_pair((C.c_uint8 * 32).from_buffer_copy(left), (C.c_uint8 * 32).from_buffer_copy(right), out)
    return bytes(out)


def merkle_build(elements, batch_size: int):
    """BinaryMerkleTree::build over B128 elements ((n, 2) uint64): list of 2 * n_leaves - 1 digests, root last"""
    e = _c(elements)
    n_leaves = len(e) // batch_size
    nodes = np.zeros((2 * n_leaves - 1) * 32, np.uint8)
    lib().orc_merkle_build(_p(e), C.c_uint64(n_leaves), C.c_uint64(16 * batch_size), _p(nodes))
    return [bytes(nodes[32 * i: 32 * i + 32]) for i in range(2 * n_leaves - 1)]


# ------------------------------------------------------------------------------------------------
# deterministic inputs: SplitMix64 (documented generator of SURVEY.md 8d; the reference's rand 0.9
# ChaCha12 StdRng is not available outside Rust and the reference stores no golden op outputs)
def splitmix64(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        out[:] = z ^ (z >> np.uint64(31))
    return out


def rand_b128(seed: int, n: int) -> np.ndarray:
    return splitmix64(seed, 2 * n).reshape(n, 2)
