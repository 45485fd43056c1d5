"""Host-side scalar tower arithmetic (python ints) for the O(1)-per-round scalar work of the host
mirror: batch-coefficient powers, eq-indicator shard factors, round-polynomial bookkeeping.
Definition: reference crates/field/src/arch/portable/pairwise_recursive_arithmetic.rs:12-62.
(Independent of oracle/, which is test infrastructure.)"""
from functools import lru_cache


def _mul_alpha(a: int, k: int) -> int:
    if k == 0:
        return a
    h = 1 << (k - 1)
    m = (1 << h) - 1
    a0, a1 = a & m, a >> h
    return a1 | ((a0 ^ _mul_alpha(a1, k - 1)) << h)


@lru_cache(maxsize=1 << 16)
def _mul8(a: int, b: int) -> int:
    return _mul(a, b, 3, False)


def _mul(a: int, b: int, k: int, use_cache: bool = True) -> int:
    if a == 0 or b == 0:
        return 0
    if k == 0:
        return a & b
    if k == 3 and use_cache:
        return _mul8(a, b)
    h = 1 << (k - 1)
    m = (1 << h) - 1
    a0, a1, b0, b1 = a & m, a >> h, b & m, b >> h
    z0, z2 = _mul(a0, b0, k - 1), _mul(a1, b1, k - 1)
    z1 = _mul(a0 ^ a1, b0 ^ b1, k - 1) ^ z0 ^ z2
    return (z0 ^ z2) | ((z1 ^ _mul_alpha(z2, k - 1)) << h)


_M64 = (1 << 64) - 1


def mul(a: int, b: int, k: int = 7) -> int:
    """product in T_k (k = 7: BinaryField128b).  B128 products go through the library's host-side helper
    (b200_host_mul128, ~1 us) when it is loaded; the pure-Python recursion is the definition and the fallback."""
    if k == 7:
        try:
            import ctypes as C

            from . import _lib

            lib = _lib.load()
            A2 = C.c_uint64 * 2
            out = A2()
            lib.b200_host_mul128(A2(a & _M64, a >> 64), A2(b & _M64, b >> 64), out)
            return int(out[0]) | (int(out[1]) << 64)
        except ImportError:
            pass
    return _mul(a, b, k)


def pow_(a: int, e: int, k: int = 7) -> int:
    r = 1
    while e:
        if e & 1:
            r = mul(r, a, k)
        a = mul(a, a, k)
        e >>= 1
    return r


def eq_ind_scalar(bits: int, point) -> int:
    """prod_k (r_k if bit k of `bits` else 1 - r_k): the eq-indicator at a boolean point"""
    e = 1
    for k, r in enumerate(point):
        e = mul(e, r if (bits >> k) & 1 else r ^ 1)
    return e
