"""Pin the oracle's tower-field arithmetic against every field KAT the reference holds
(SURVEY.md 8c): mul vectors, generator orders, tower<->AES isomorphism, POLYVAL KATs through the
reference's 128x128 basis-change tables.  CPU only."""
import random

import pytest

K_OF_BITS = {1: 0, 2: 1, 4: 2, 8: 3, 16: 4, 32: 5, 64: 6, 128: 7}


@pytest.mark.parametrize("bits", [2, 4, 8, 16, 64])
def test_reference_mul_kats(oracle, kat, bits):
    # crates/field/src/binary_field.rs:925-1028
    for a, b, c in kat["mul_kats"][str(bits)]:
        assert oracle.mul(a, b, K_OF_BITS[bits]) == c
        assert oracle.mul_slow(a, b, K_OF_BITS[bits]) == c
        assert oracle.mul(b, a, K_OF_BITS[bits]) == c


def _factor(n):
    fs, p = [], 2
    while p * p <= n:
        while n % p == 0:
            fs.append(p)
            n //= p
        p += 1 if p == 2 else 2
    if n > 1:
        fs.append(n)
    return fs


@pytest.mark.parametrize("bits", [1, 2, 4, 8, 16, 32, 64, 128])
def test_multiplicative_generators(oracle, kat, bits):
    # binary_field.rs:740-747 + order test :1031-1102.  2^128-1 = prod of Fermat numbers F0..F6
    k = K_OF_BITS[bits]
    g = kat["generators"][str(bits)]
    order = (1 << bits) - 1
    if bits == 128:
        primes = [3, 5, 17, 257, 65537, 641, 6700417, 274177, 67280421310721]
    else:
        primes = sorted(set(_factor(order)))
    prod = 1
    for p in primes:
        assert order % p == 0
        prod *= p
    assert prod == order or bits == 1  # all Fermat factors are distinct (square-free order)
    assert oracle.pow_(g, order, k) == 1
    for p in primes:
        assert oracle.pow_(g, order // p, k) != 1, f"generator of B{bits} has order dividing (2^{bits}-1)/{p}"


def test_fast_mul_matches_bit_recursion(oracle):
    rng = random.Random(0)
    for k in range(8):
        bits = 1 << k
        for _ in range(200):
            a, b = rng.getrandbits(bits), rng.getrandbits(bits)
            assert oracle.mul(a, b, k) == oracle.mul_slow(a, b, k)


def test_field_axioms_and_subfield_embedding(oracle):
    rng = random.Random(1)
    for k in range(1, 8):
        bits = 1 << k
        for _ in range(50):
            a, b, c = (rng.getrandbits(bits) for _ in range(3))
            assert oracle.mul(a, oracle.mul(b, c, k), k) == oracle.mul(oracle.mul(a, b, k), c, k)
            assert oracle.mul(a, b ^ c, k) == oracle.mul(a, b, k) ^ oracle.mul(a, c, k)
            assert oracle.mul(a, 1, k) == a
            assert oracle.square(a, k) == oracle.mul(a, a, k)
            if a:
                assert oracle.mul(a, oracle.invert(a, k), k) == 1
            # subfield = high half zero (binary_field.rs:505-527): product stays in the subfield
            lo_a, lo_b = a & ((1 << (bits // 2)) - 1), b & ((1 << (bits // 2)) - 1)
            assert oracle.mul(lo_a, lo_b, k) == oracle.mul(lo_a, lo_b, k - 1)
        assert oracle.invert(0, k) == 0
        # mul_alpha multiplies by X_{k-1} = 1 << 2^(k-1)
        for _ in range(20):
            a = rng.getrandbits(bits)
            assert oracle.mul_alpha(a, k) == oracle.mul(a, 1 << (bits // 2), k)


def test_mul_by_subfield_is_limbwise(oracle):
    # binary_field.rs:363-414 ; limb layout KAT :1155-1169 (from_bases: low limb first)
    rng = random.Random(2)
    for k in range(0, 8):
        w = 1 << k
        for _ in range(30):
            a, s = rng.getrandbits(128), rng.getrandbits(w)
            assert oracle.mul_subfield(a, s, k) == oracle.mul(a, s, 7)
    assert oracle.mul_subfield(0x04030201, 1, 3) == 0x04030201


def _lin(table, v):
    r, i = 0, 0
    while v:
        if v & 1:
            r ^= table[i]
        v >>= 1
        i += 1
    return r


def _aes_mul(a, b):
    r = 0
    while b:
        if b & 1:
            r ^= a
        a <<= 1
        if a & 0x100:
            a ^= 0x11B
        b >>= 1
    return r


def test_tower_aes_isomorphism(oracle, kat):
    # aes_field.rs:113-141 : byte-wise basis change commutes with multiplication (GF(256) mod 0x11B)
    t2a, a2t = kat["binary_to_aes"], kat["aes_to_binary"]
    for v in range(256):
        assert _lin(a2t, _lin(t2a, v)) == v
    for a in range(256):
        for b in range(0, 256, 7):
            assert _lin(t2a, oracle.mul(a, b, 3)) == _aes_mul(_lin(t2a, a), _lin(t2a, b))
    # level-3 alpha (0x10 in the tower) maps to 0xd3 (aes_field.rs:230-235)
    assert _lin(t2a, 0x10) == 0xD3
    assert _lin(t2a, kat["generators"]["8"]) == kat["aes_generators"]["8"]


POLY = (1 << 128) | (1 << 127) | (1 << 126) | (1 << 121) | 1


def _clmul_mod(a, b):
    r = 0
    while b:
        if b & 1:
            r ^= a
        b >>= 1
        a <<= 1
        if a >> 128:
            a ^= POLY
    return r


def _pv_pow(a, e):
    r = 1
    while e:
        if e & 1:
            r = _clmul_mod(r, a)
        a = _clmul_mod(a, a)
        e >>= 1
    return r


def test_polyval_kats_pin_b128_mul(oracle, kat):
    # polyval.rs:262 ONE, :308-311 to_montgomery, :1113-1127 mul/sqr KATs, :516-788 tables,
    # :1150-1156 "conversion commutes with multiplication"
    x128 = _clmul_mod(1 << 127, 2)
    inv_r = _pv_pow(x128, (1 << 128) - 2)

    def mont(a, b):
        return _clmul_mod(_clmul_mod(a, b), inv_r)

    one = int(kat["polyval_one"], 16)
    assert one == x128  # Montgomery form of 1 is x^128 mod p
    r2 = int(kat["polyval_to_montgomery_const"], 16)
    assert r2 == _clmul_mod(x128, x128) or mont(1, r2) == one

    def new(v):
        return mont(v, r2)

    a, b, c = (int(x, 16) for x in kat["polyval_mul_kat"])
    assert mont(new(a), new(b)) == new(c)
    sa, sc = (int(x, 16) for x in kat["polyval_sqr_kat"])
    assert mont(new(sa), new(sa)) == new(sc)

    b2p = [int(x, 16) for x in kat["binary_to_polyval"]]
    p2b = [int(x, 16) for x in kat["polyval_to_binary"]]
    assert _lin(b2p, 1) == one
    rng = random.Random(3)
    for _ in range(64):
        u, v = rng.getrandbits(128), rng.getrandbits(128)
        assert _lin(p2b, _lin(b2p, u)) == u
        assert _lin(p2b, mont(_lin(b2p, u), _lin(b2p, v))) == oracle.mul(u, v, 7)
    # the POLYVAL KAT itself, pulled back to the tower: tower(a)*tower(b) == tower(c)
    ta, tb, tc = (_lin(p2b, new(x)) for x in (a, b, c))
    assert oracle.mul(ta, tb, 7) == tc


def test_generator_maps_to_polyval_generator(oracle, kat):
    # binary_field.rs:747 and polyval.rs:496 are related by the basis change (polyval.rs:516-648)
    b2p = [int(x, 16) for x in kat["binary_to_polyval"]]
    assert _lin(b2p, kat["generators"]["128"]) == int(kat["polyval_generator"], 16)
