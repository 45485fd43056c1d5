"""CPU checks of the POLYVAL pieces: the oracle's Montgomery multiply (oracle/polyval.c,
restating arch/portable/packed_polyval_128.rs:88-160) against the reference's KATs (polyval.rs:1113-1127), the
library's own bit-serial host product against it, and the library's DERIVED tower <-> POLYVAL basis change against the
reference's tables (polyval.rs:516-788, extracted into tests/golden/field_kat.json)."""
import random


def test_oracle_polyval_mul_kats(oracle, kat):
    r2 = int(kat["polyval_to_montgomery_const"], 16)
    new = lambda v: oracle.polyval_mul(v, r2)  # noqa: E731  BinaryField128bPolyval::new = to_montgomery (polyval.rs:56-60, 308-311)
    a, b, c = (int(x, 16) for x in kat["polyval_mul_kat"])
    assert oracle.polyval_mul(new(a), new(b)) == new(c)
    sa, sc = (int(x, 16) for x in kat["polyval_sqr_kat"])
    assert oracle.polyval_mul(new(sa), new(sa)) == new(sc)
    one = int(kat["polyval_one"], 16)
    assert new(1) == one and oracle.polyval_mul(one, new(a)) == new(a)


def test_derived_basis_change_equals_reference_tables(oracle, kat):
    from binius_b200 import polyval as pv

    t2p, p2t = pv.basis_change()
    assert [hex(x) for x in t2p] == kat["binary_to_polyval"]
    assert [hex(x) for x in p2t] == kat["polyval_to_binary"]
    assert pv.to_polyval(1) == pv.ONE == int(kat["polyval_one"], 16)
    assert pv.to_polyval(int(kat["generators"]["128"], 16) if isinstance(kat["generators"]["128"], str) else kat["generators"]["128"]) == int(kat["polyval_generator"], 16)


def test_host_product_and_isomorphism(oracle):
    from binius_b200 import polyval as pv

    rng = random.Random(3)
    for _ in range(300):
        x, y = rng.getrandbits(128), rng.getrandbits(128)
        assert pv.mul(x, y) == oracle.polyval_mul(x, y)
        assert pv.to_tower(pv.to_polyval(x)) == x
        # conversion commutes with multiplication (polyval.rs:1150-1156)
        assert pv.to_polyval(oracle.mul(x, y)) == oracle.polyval_mul(pv.to_polyval(x), pv.to_polyval(y))


def test_oracle_gpa_layers_and_linear_map(oracle):
    from binius_b200 import polyval as pv

    n_vars = 6
    x = oracle.rand_b128(41, 1 << n_vars)
    t2p, p2t = pv.basis_change()
    xp = oracle.linear_map(t2p, x)
    assert oracle.to_ints(xp) == [pv.to_polyval(v) for v in oracle.to_ints(x)]
    layers = oracle.polyval_gpa_layers(xp, n_vars)
    assert [len(l) for l in layers] == [1 << (n_vars - k) for k in range(n_vars + 1)]
    # the grand product equals the image of the tower product of all inputs
    prod = 1
    for v in oracle.to_ints(x):
        prod = oracle.mul(prod, v)
    assert oracle.to_ints(layers[-1])[0] == pv.to_polyval(prod)
