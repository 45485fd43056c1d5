"""Host-side mirror of `binius_ntt::AdditiveNTT<F>` (reference crates/ntt/src/additive_ntt.rs:58-166)
over the C ABI.  `NTTShape` = additive_ntt.rs:20-56.  Data of the trait methods is a HOST array (the
trait takes `&mut [P]`); the `*_device` variants work on device-resident slices (what the commit
pipeline uses once the codeword lives on the GPU)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .layer import B200Layer, DevSlice, InputValidation

_DT = {3: np.uint8, 4: np.uint16, 5: np.uint32, 6: np.uint64}


@dataclass
class NTTShape:
    log_x: int = 0
    log_y: int = 0
    log_z: int = 0


class B200AdditiveNTT:
    """SingleThreadedNTT::new(log_domain_size) equivalent (single_threaded.rs:27-45) for
    F = BinaryField{8,16,32}b (`field_log_bits` = 3, 4, 5), subspace <1, 2, 4, ...>."""

    def __init__(self, layer: B200Layer, field_log_bits: int, log_domain_size: int):
        self._l = layer
        self.field_log_bits = field_log_bits
        h = C.c_void_p()
        layer._check(layer._lib.b200_ntt_create(layer._ctx, field_log_bits, log_domain_size, C.byref(h)))
        self.handle = h

    def log_domain_size(self) -> int:
        return int(self._l._lib.b200_ntt_log_domain_size(self.handle))

    def get_subspace_eval(self, i: int, j: int) -> int:
        out = (C.c_uint64 * 2)()
        rc = self._l._lib.b200_ntt_get_subspace_eval(self.handle, i, j, out)
        if rc:
            raise InputValidation(f"get_subspace_eval({i}, {j}) out of range")
        return int(out[0]) | (int(out[1]) << 64)

    # -- host data (trait signature) ----------------------------------------------------------------
    def _host(self, fn, data: np.ndarray, shape: NTTShape, coset, coset_bits, skip_rounds):
        kd, n = _elem_info(data)
        assert data.flags["C_CONTIGUOUS"]
        self._l._check(fn(self._l._ctx, self.handle, data.ctypes.data, kd, n, shape.log_x, shape.log_y, shape.log_z,
                          coset, coset_bits, skip_rounds))

    def forward_transform(self, data: np.ndarray, shape: NTTShape, coset: int = 0, coset_bits: int = 0, skip_rounds: int = 0):
        self._host(self._l._lib.b200_ntt_forward_host, data, shape, coset, coset_bits, skip_rounds)

    def inverse_transform(self, data: np.ndarray, shape: NTTShape, coset: int = 0, coset_bits: int = 0, skip_rounds: int = 0):
        self._host(self._l._lib.b200_ntt_inverse_host, data, shape, coset, coset_bits, skip_rounds)

    # *_transform_ext (additive_ntt.rs:137-165): data = packed extension elements ((n,2) uint64 = B128)
    forward_transform_ext = forward_transform
    inverse_transform_ext = inverse_transform

    # -- device data ----------------------------------------------------------------------------------
    def forward_device(self, ptr: int, elem_log_bits: int, n_elems: int, shape: NTTShape, coset=0, coset_bits=0, skip_rounds=0):
        self._l._check(self._l._lib.b200_ntt_forward(self._l._ctx, self.handle, ptr, elem_log_bits, n_elems, shape.log_x,
                                                     shape.log_y, shape.log_z, coset, coset_bits, skip_rounds))

    def inverse_device(self, ptr: int, elem_log_bits: int, n_elems: int, shape: NTTShape, coset=0, coset_bits=0, skip_rounds=0):
        self._l._check(self._l._lib.b200_ntt_inverse(self._l._ctx, self.handle, ptr, elem_log_bits, n_elems, shape.log_x,
                                                     shape.log_y, shape.log_z, coset, coset_bits, skip_rounds))

    def __del__(self):
        try:
            self._l._lib.b200_ntt_destroy(self.handle)
        except Exception:
            pass


def _elem_info(data: np.ndarray):
    if data.dtype == np.uint64 and data.ndim == 2 and data.shape[1] == 2:
        return 7, data.shape[0]
    for k, dt in _DT.items():
        if data.dtype == dt and data.ndim == 1:
            return k, data.shape[0]
    raise InputValidation(f"unsupported element array dtype={data.dtype} shape={data.shape}")
