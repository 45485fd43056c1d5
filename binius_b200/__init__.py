"""binius_b200 -- B200-native (sm_100a) implementation of the Binius prover hot path behind the
reference's ComputeLayer / AdditiveNTT / ComputationBackend surfaces.

The package is a thin host-side mirror of those interfaces over the C ABI in
include/binius_b200.h; all arithmetic runs in hand-written CUDA kernels (binius_b200/csrc).
There is no CPU fallback: importing works without a GPU (so the ABI can be inspected), but creating
a `B200Layer` without a usable sm_100 device raises `DeviceError`.
"""
from . import _lib  # noqa: F401
from .layer import (AllocError, ArithCircuit, B200Executor, B200KernelExecutor, B200Layer, B200LayerHolder,  # noqa: F401
                    BumpAllocator, ComputeData, DevSlice, DeviceError, Error, ExprEval, HostBumpAllocator,
                    InputValidation, KernelBuffer, KernelMemMap, NttError, OpValue, SlicesBatch, SubfieldSlice,
                    calculate_round_evals, eq_ind_partial_eval, to_arr, to_ints)
from .ntt import B200AdditiveNTT, NTTShape  # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
