"""ctypes loader for the sm_100a shared library (C ABI: include/binius_b200.h).

There is no CPU fallback: if the library is missing or no B200 is usable this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbinius_b200.so")

OK, ERR_INPUT_VALIDATION, ERR_ALLOC, ERR_DEVICE = 0, 1, 2, 3
ERR_NTT_POWER_OF_TWO, ERR_NTT_SKIP_ROUNDS, ERR_NTT_BATCH, ERR_NTT_COSET, ERR_NTT_DOMAIN, ERR_NTT_FIELD = 11, 12, 13, 14, 15, 16

# every symbol include/binius_b200.h declares (checked by tests/test_abi.py against the header)
SYMBOLS = [
    "b200_ctx_create", "b200_ctx_destroy", "b200_last_error", "b200_ctx_stream", "b200_flush", "b200_ctx_set_stream", "b200_ctx_set_tuning",
    "b200_kernel_scope_begin", "b200_kernel_local", "b200_kernel_scope_end", "b200_host_mul128", "b200_sumcheck_tail_start", "b200_sumcheck_tail_round_evals", "b200_sumcheck_tail_challenge",
    "b200_sumcheck_tail_finish", "b200_groestl256_leaves", "b200_groestl256_compress_pairs", "b200_merkle_build", "b200_linear_map", "b200_host_polyval_mul", "b200_host_polyval_basis_change",
    "b200_event_create", "b200_event_record", "b200_event_elapsed_ms", "b200_event_destroy",
    "b200_ctx_launch_count", "b200_dev_alloc", "b200_dev_free", "b200_host_alloc", "b200_host_free",
    "b200_copy_h2d", "b200_copy_h2d_side", "b200_side_join", "b200_copy_d2h", "b200_copy_d2d", "b200_fill", "b200_sync", "b200_results_reset",
    "b200_results_fetch", "b200_extrapolate_line", "b200_extrapolate_line_host", "b200_tensor_expand", "b200_inner_product", "b200_fold_left",
    "b200_fold_right", "b200_expr_compile", "b200_expr_free", "b200_expr_n_vars", "b200_compute_composite",
    "b200_pairwise_product_reduce", "b200_kernel_decl_value", "b200_kernel_sum_composition_evals",
    "b200_kernel_add", "b200_kernel_add_assign", "b200_bivariate_round_evals", "b200_ntt_create",
    "b200_ntt_destroy", "b200_ntt_log_domain_size", "b200_ntt_get_subspace_eval", "b200_ntt_forward",
    "b200_ntt_inverse", "b200_ntt_forward_host", "b200_ntt_inverse_host", "b200_fri_fold",
    "b200_tensor_product_full_query", "b200_fold_multilinears_high_to_low", "b200_eq_ind_round_evals",
    "b200_sumcheck_round_evals", "b200_fold_multilinears_low_to_high", "b200_zerocheck_univariate_evals", "b200_zerocheck_univariate_evals_streamed",
    "b200_zerocheck_univariate_store_elems", "b200_zerocheck_univariate_prepare", "b200_zerocheck_univariate_finish",
]


class ExprStep(C.Structure):
    _fields_ = [("op", C.c_uint32), ("l", C.c_uint32), ("r", C.c_uint64), ("c_lo", C.c_uint64), ("c_hi", C.c_uint64)]


_lib = None


def load() -> C.CDLL:
    """Load the shared library (no device needed). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make` or `python -c 'import __graft_entry__ as g; g.build()'`. "
            "binius_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
    P = C.POINTER
    sig = {
        "b200_ctx_create": (i32, [i32, P(vp)]),
        "b200_ctx_destroy": (None, [vp]),
        "b200_last_error": (C.c_char_p, [vp]),
        "b200_ctx_stream": (vp, [vp]),
        "b200_ctx_set_stream": (i32, [vp, vp]),
        "b200_ctx_set_tuning": (i32, [vp, C.c_char_p, i32]),
        "b200_flush": (i32, [vp]),
        "b200_host_mul128": (None, [P(u64), P(u64), P(u64)]),
        "b200_linear_map": (i32, [vp, vp, vp, u64, P(u64)]),
        "b200_sumcheck_tail_start": (i32, [vp, P(vp), u32, u32, vp, P(vp), P(vp), u32, P(u32), P(u64), u32, u32, P(vp)]),
        "b200_sumcheck_tail_round_evals": (i32, [vp, P(u64)]),
        "b200_sumcheck_tail_challenge": (i32, [vp, P(u64)]),
        "b200_sumcheck_tail_finish": (i32, [vp]),
        "b200_groestl256_leaves": (i32, [vp, vp, u64, u64, vp]),
        "b200_groestl256_compress_pairs": (i32, [vp, vp, u64, vp]),
        "b200_merkle_build": (i32, [vp, vp, u64, u64, vp, u64]),
        "b200_host_polyval_mul": (None, [P(u64), P(u64), P(u64)]),
        "b200_host_polyval_basis_change": (i32, [P(u64), P(u64)]),
        "b200_kernel_scope_begin": (i32, [vp]),
        "b200_kernel_local": (i32, [vp, u32, P(vp)]),
        "b200_kernel_scope_end": (i32, [vp]),
        "b200_event_create": (i32, [vp, P(vp)]),
        "b200_event_record": (i32, [vp, vp]),
        "b200_event_elapsed_ms": (i32, [vp, vp, vp, P(C.c_float)]),
        "b200_event_destroy": (None, [vp]),
        "b200_ctx_launch_count": (u64, [vp]),
        "b200_dev_alloc": (i32, [vp, u64, P(vp)]),
        "b200_dev_free": (i32, [vp, vp]),
        "b200_host_alloc": (i32, [vp, u64, P(vp)]),
        "b200_host_free": (i32, [vp, vp]),
        "b200_copy_h2d": (i32, [vp, vp, vp, u64]),
        "b200_copy_d2h": (i32, [vp, vp, vp, u64]),
        "b200_copy_h2d_side": (i32, [vp, vp, vp, u64]),
        "b200_side_join": (i32, [vp]),
        "b200_copy_d2d": (i32, [vp, vp, vp, u64]),
        "b200_fill": (i32, [vp, vp, u64, P(u64)]),
        "b200_sync": (i32, [vp]),
        "b200_results_reset": (i32, [vp]),
        "b200_results_fetch": (i32, [vp, P(u32), u32, P(u64)]),
        "b200_extrapolate_line": (i32, [vp, vp, u64, vp, u64, P(u64)]),
        "b200_extrapolate_line_host": (i32, [vp, vp, vp, u64, P(u64)]),
        "b200_tensor_expand": (i32, [vp, vp, u64, u32, P(u64), u32]),
        "b200_inner_product": (i32, [vp, vp, u64, u32, vp, u64, P(u32)]),
        "b200_fold_left": (i32, [vp, vp, u64, u32, vp, u64, vp, u64]),
        "b200_fold_right": (i32, [vp, vp, u64, u32, vp, u64, vp, u64]),
        "b200_expr_compile": (i32, [vp, P(ExprStep), u32, P(vp)]),
        "b200_expr_free": (None, [vp]),
        "b200_expr_n_vars": (u32, [vp]),
        "b200_compute_composite": (i32, [vp, P(vp), u32, u64, vp, u64, vp]),
        "b200_pairwise_product_reduce": (i32, [vp, vp, u64, P(vp), P(u64), u32]),
        "b200_kernel_decl_value": (i32, [vp, P(u64), P(u32)]),
        "b200_kernel_sum_composition_evals": (i32, [vp, P(vp), u32, u64, vp, P(u64), u32]),
        "b200_kernel_add": (i32, [vp, u32, vp, vp, vp]),
        "b200_kernel_add_assign": (i32, [vp, u32, vp, vp]),
        "b200_bivariate_round_evals": (i32, [vp, P(vp), u32, u32, P(u32), P(u32), u32, P(u64), P(u32), P(u32)]),
        "b200_ntt_create": (i32, [vp, u32, u32, P(vp)]),
        "b200_ntt_destroy": (None, [vp]),
        "b200_ntt_log_domain_size": (u32, [vp]),
        "b200_ntt_get_subspace_eval": (i32, [vp, u32, u64, P(u64)]),
        "b200_ntt_forward": (i32, [vp, vp, vp, u32, u64, u32, u32, u32, u64, u32, u32]),
        "b200_ntt_inverse": (i32, [vp, vp, vp, u32, u64, u32, u32, u32, u64, u32, u32]),
        "b200_ntt_forward_host": (i32, [vp, vp, vp, u32, u64, u32, u32, u32, u64, u32, u32]),
        "b200_ntt_inverse_host": (i32, [vp, vp, vp, u32, u64, u32, u32, u32, u64, u32, u32]),
        "b200_fri_fold": (i32, [vp, vp, u32, u32, P(u64), u32, vp, u64, vp, u64]),
        "b200_tensor_product_full_query": (i32, [vp, P(u64), u32, vp, u64]),
        "b200_fold_multilinears_high_to_low": (i32, [vp, P(vp), u32, u32, P(u64), P(u64), P(u64), P(u64)]),
        "b200_eq_ind_round_evals": (i32, [vp, P(vp), P(u64), P(u64), u32, u32, vp, P(vp), P(vp), u32, P(u32), P(u64), u32, P(u32)]),
        "b200_sumcheck_round_evals": (i32, [vp, u32, P(vp), P(u64), P(u64), u32, u32, vp, P(vp), P(vp), u32, P(u32), P(u64), u32, P(u32)]),
        "b200_fold_multilinears_low_to_high": (i32, [vp, P(vp), P(vp), u32, u32, P(u64), P(u64), P(u64), P(u64)]),
        "b200_zerocheck_univariate_evals": (i32, [vp, P(vp), P(u32), u32, u32, u32, vp, u64, P(vp), P(u32), u32, u32, P(u64)]),
        "b200_zerocheck_univariate_evals_streamed": (i32, [vp, P(vp), P(vp), P(u32), u32, u32, u32, vp, u64, P(vp), P(u32), u32, u32, u32, P(u64)]),
        "b200_zerocheck_univariate_store_elems": (u64, [u32, u32, P(u32), u32]),
        "b200_zerocheck_univariate_prepare": (i32, [vp, P(vp), P(vp), P(u32), u32, u32, u32, P(vp), P(u32), u32, u32, u32, vp, u64, P(u32)]),
        "b200_zerocheck_univariate_finish": (i32, [vp, P(vp), P(u32), u32, u32, u32, vp, u64, P(vp), P(u32), u32, u32, vp, u64, P(u64)]),
    }
    for name in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the build is stale: fail loudly
        fn.restype, fn.argtypes = sig[name]
    _lib = lib
    return lib
