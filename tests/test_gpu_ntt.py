"""GPU parity tests of the additive NTT (AdditiveNTT trait surface through the C ABI) vs the CPU
oracle; mirrors crates/ntt/src/tests/ntt_tests.rs (all shapes / cosets / skip_rounds, round trip)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hal():
    import binius_b200

    layer = binius_b200.B200Layer(0)
    yield layer
    layer.close()


SHAPES = [  # log_x, log_y, log_z, coset_bits, skip
    (0, 5, 0, 0, 0), (2, 4, 1, 2, 0), (3, 3, 0, 1, 1), (0, 6, 2, 0, 2), (1, 5, 0, 3, 5), (0, 1, 0, 0, 0),
    (0, 12, 0, 0, 0), (6, 9, 0, 1, 1), (0, 14, 1, 2, 0), (3, 11, 2, 0, 3), (5, 10, 0, 0, 0), (0, 16, 0, 0, 0),
]


@pytest.mark.parametrize("kt,dt", [(5, np.uint32), (4, np.uint16), (3, np.uint8)])
def test_ntt_matches_oracle_all_shapes(hal, oracle, kt, dt):
    import binius_b200

    d = {5: 20, 4: 16, 3: 8}[kt]
    ntt = binius_b200.B200AdditiveNTT(hal, kt, d)
    ontt = oracle.NTT(kt, d)
    assert ntt.log_domain_size() == d
    rng = np.random.default_rng(kt)
    for (lx, ly, lz, cb, skip) in SHAPES:
        if ly + cb > d:
            continue
        n = 1 << (lx + ly + lz)
        data = rng.integers(0, 1 << (1 << kt), size=n, dtype=np.uint64).astype(dt)
        coset = (1 << cb) - 1
        shape = binius_b200.NTTShape(lx, ly, lz)
        f = data.copy()
        ntt.forward_transform(f, shape, coset, cb, skip)
        assert np.array_equal(f, ontt.forward(data, kt, lx, ly, lz, coset, cb, skip)), (lx, ly, lz, cb, skip)
        b = f.copy()
        ntt.inverse_transform(b, shape, coset, cb, skip)
        assert np.array_equal(b, data)
        inv = data.copy()
        ntt.inverse_transform(inv, shape, coset, cb, skip)
        assert np.array_equal(inv, ontt.inverse(data, kt, lx, ly, lz, coset, cb, skip))


def test_ntt_ext_b128(hal, oracle):
    # additive_ntt.rs:137-165: packed B128 data with B32 twiddles
    import binius_b200

    ntt = binius_b200.B200AdditiveNTT(hal, 5, 16)
    ontt = oracle.NTT(5, 16)
    for (lx, ly, lz, cb, skip) in [(0, 6, 0, 0, 0), (2, 10, 0, 1, 1), (4, 8, 1, 0, 0)]:
        data = oracle.rand_b128(ly, 1 << (lx + ly + lz))
        f = data.copy()
        ntt.forward_transform_ext(f, binius_b200.NTTShape(lx, ly, lz), (1 << cb) - 1, cb, skip)
        assert np.array_equal(f, ontt.forward(data, 7, lx, ly, lz, (1 << cb) - 1, cb, skip))
        ntt.inverse_transform_ext(f, binius_b200.NTTShape(lx, ly, lz), (1 << cb) - 1, cb, skip)
        assert np.array_equal(f, data)


def test_ntt_errors(hal):
    # crates/ntt/src/single_threaded.rs:364-406
    import binius_b200

    ntt = binius_b200.B200AdditiveNTT(hal, 5, 10)
    data = np.zeros(64, dtype=np.uint32)
    S = binius_b200.NTTShape
    for args, kind in [((S(0, 6, 0), 0, 0, 7), "SkipRoundsTooLarge"), ((S(0, 5, 0), 0, 0, 0), "BatchTooLarge"),
                       ((S(0, 6, 0), 4, 2, 0), "CosetIndexOutOfBounds"), ((S(0, 6, 0), 0, 5, 0), "DomainTooSmall")]:
        with pytest.raises(binius_b200.NttError) as e:
            ntt.forward_transform(data, *args)
        assert e.value.kind == kind
    with pytest.raises(binius_b200.NttError) as e:
        ntt.forward_transform(np.zeros(48, dtype=np.uint32), S(0, 5, 0))
    assert e.value.kind == "PowerOfTwoLengthRequired"
    with pytest.raises(binius_b200.NttError) as e:
        binius_b200.B200AdditiveNTT(hal, 3, 9)
    assert e.value.kind == "FieldTooSmall"


def test_ntt_large_roundtrip_and_linearity(hal, oracle):
    # BASELINE cfg#2 size (2^24 B32 coefficients): size-independent properties + oracle spot check on
    # a sub-transform (the batched layout makes every z-slab an independent transform)
    import binius_b200

    ntt = binius_b200.B200AdditiveNTT(hal, 5, 24)
    rng = np.random.default_rng(0)
    for (lx, ly, lz, skip) in [(6, 18, 0, 1), (0, 24, 0, 0), (0, 16, 8, 0)]:
        n = 1 << 24
        a = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
        b = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
        S = binius_b200.NTTShape(lx, ly, lz)
        fa, fb, fab = a.copy(), b.copy(), a ^ b
        ntt.forward_transform(fa, S, 0, 0, skip)
        ntt.forward_transform(fb, S, 0, 0, skip)
        ntt.forward_transform(fab, S, 0, 0, skip)
        assert np.array_equal(fab, fa ^ fb)
        back = fa.copy()
        ntt.inverse_transform(back, S, 0, 0, skip)
        assert np.array_equal(back, a)
    # oracle spot check: S3 slab 5 equals a standalone 2^16 transform
    ontt = oracle.NTT(5, 24)
    slab = a[5 << 16: 6 << 16]
    assert np.array_equal(fa[5 << 16: 6 << 16], ontt.forward(slab, 5, 0, 16, 0))
