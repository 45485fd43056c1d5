"""Seeded randomised differential tests: CUDA path (through the C ABI) vs the CPU oracle on random shapes,
to exercise the size-dependent kernel selection (small / fused / TMA / tensor-core paths, ragged tiles,
unaligned device slices).  Bit-exact."""
import ctypes as C
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hal():
    import binius_b200

    layer = binius_b200.B200Layer(0)
    yield layer
    layer.close()


def _same(a, b):
    return np.array_equal(np.asarray(a, dtype=np.uint64).reshape(-1, 2), np.asarray(b, dtype=np.uint64).reshape(-1, 2))


def test_fuzz_fold_multilinears_both_orders(hal, oracle):
    rng = random.Random(20260101)
    for case in range(12):
        n_vars = rng.randint(1, 17)
        full = 1 << n_vars
        m = rng.randint(1, 7)
        prefixes = [rng.choice([full, rng.randint(0, full), max(full // 2 + rng.randint(-3, 3), 0)]) for _ in range(m)]
        prefixes = [min(max(p, 0), full) for p in prefixes]
        suffixes = [rng.choice([0, 1, rng.getrandbits(128)]) for _ in range(m)]
        z = rng.choice([0, 1, rng.getrandbits(128), rng.getrandbits(8)])
        host = [oracle.rand_b128(case * 100 + t, p) if p else np.zeros((0, 2), np.uint64) for t, p in enumerate(prefixes)]
        # device copies at an unaligned element offset inside a bigger allocation
        devs = []
        for h in host:
            d = hal.dev_alloc(len(h) + 3)
            if len(h):
                hal.copy_h2d(h, d.slice(3, 3 + len(h)))
            devs.append(d.slice(3, 3 + len(h)))
        zs = (C.c_uint64 * 2)(z & (2**64 - 1), z >> 64)
        sfx = (C.c_uint64 * (2 * m))(*[w for s in suffixes for w in (s & (2**64 - 1), s >> 64)])
        lens = (C.c_uint64 * m)(*prefixes)
        new_lens = (C.c_uint64 * m)()
        outs = [hal.dev_alloc(max((p + 1) // 2, 1)) for p in prefixes]
        ptrs = (C.c_void_p * m)(*[d.ptr for d in devs])
        optrs = (C.c_void_p * m)(*[o.ptr for o in outs])
        hal._check(hal._lib.b200_fold_multilinears_low_to_high(hal._ctx, ptrs, optrs, m, n_vars, lens, sfx, zs, new_lens))
        for t in range(m):
            exp = oracle.fold_right_lerp(host[t], suffixes[t], z)
            assert int(new_lens[t]) == len(exp)
            if len(exp):
                assert _same(hal.to_host(outs[t].slice(0, len(exp))), exp), (case, t, "low_to_high")
        hal._check(hal._lib.b200_fold_multilinears_high_to_low(hal._ctx, ptrs, m, n_vars, lens, sfx, zs, new_lens))
        for t in range(m):
            exp = oracle.fold_left_lerp_inplace(host[t] if len(host[t]) else np.zeros((1, 2), np.uint64), prefixes[t], suffixes[t], n_vars, z)
            assert int(new_lens[t]) == len(exp)
            if len(exp):
                assert _same(hal.to_host(devs[t].slice(0, len(exp))), exp), (case, t, "high_to_low")


def test_fuzz_tensor_expand_and_fri(hal, oracle):
    import binius_b200

    rng = random.Random(77)
    for case in range(10):
        log_n = rng.randint(0, 6)
        k = rng.randint(0, 16 - log_n)
        data = oracle.rand_b128(3000 + case, 1 << log_n)
        buf = np.zeros((1 << (log_n + k), 2), np.uint64)
        buf[: 1 << log_n] = data
        coords = [rng.choice([rng.getrandbits(128), 0, 1]) for _ in range(k)]
        d = hal.to_device(buf)
        hal.execute(lambda ex: (ex.tensor_expand(log_n, coords, d), [])[1])
        assert _same(hal.to_host(d), oracle.tensor_expand(buf, log_n, coords)), (case, log_n, k)
    ntt = binius_b200.B200AdditiveNTT(hal, 5, 16)
    ontt = oracle.NTT(5, 16)
    for case in range(10):
        n_ch = rng.randint(0, 5)
        log_batch = rng.randint(0, n_ch)
        eta = n_ch - log_batch
        log_len = rng.randint(max(eta, 1), 14)
        ch = [rng.getrandbits(128) for _ in range(n_ch)]
        data = oracle.rand_b128(4000 + case, 1 << (log_len + log_batch))
        n_out = 1 << (log_len - eta)
        din, dout = hal.to_device(data), hal.dev_alloc(n_out)
        hal.execute(lambda ex: (ex.fri_fold(ntt, log_len, log_batch, ch, din, dout), [])[1])
        assert _same(hal.to_host(dout), ontt.fri_fold(log_len, log_batch, ch, data, n_out)), (case, log_len, log_batch, n_ch)


def test_fuzz_ntt_shapes(hal, oracle):
    import binius_b200
    from binius_b200 import NTTShape

    rng = random.Random(5)
    ntt = binius_b200.B200AdditiveNTT(hal, 5, 20)
    ontt = oracle.NTT(5, 20)
    for case in range(10):
        log_y = rng.randint(1, 14)
        log_x = rng.randint(0, 7)
        log_z = rng.randint(0, 3)
        coset_bits = rng.randint(0, min(3, 20 - log_y))
        coset = rng.randrange(1 << coset_bits)
        skip = rng.randint(0, min(2, log_y))
        n = 1 << (log_x + log_y + log_z)
        data = oracle.splitmix64(6000 + case, n).astype(np.uint32)
        f = data.copy()
        ntt.forward_transform(f, NTTShape(log_x, log_y, log_z), coset, coset_bits, skip)
        assert np.array_equal(f, ontt.forward(data, 5, log_x, log_y, log_z, coset, coset_bits, skip)), (case, log_x, log_y, log_z, coset, skip)
        ntt.inverse_transform(f, NTTShape(log_x, log_y, log_z), coset, coset_bits, skip)
        assert np.array_equal(f, data), (case, "round trip")


def test_fold_many_multilinears_one_launch(hal, oracle):
    """More multilinears than fit the kernel-parameter segment list (48): the list is staged in device
    memory and the whole fold is one launch."""
    rng = random.Random(99)
    n_vars, m = 11, 130
    full = 1 << n_vars
    prefixes = [rng.choice([full, full, rng.randint(1, full)]) for _ in range(m)]
    suffixes = [rng.choice([0, rng.getrandbits(128)]) for _ in range(m)]
    z = rng.getrandbits(128)
    host = [oracle.rand_b128(7000 + t, p) for t, p in enumerate(prefixes)]
    devs = [hal.to_device(h) for h in host]
    zs = (C.c_uint64 * 2)(z & (2**64 - 1), z >> 64)
    sfx = (C.c_uint64 * (2 * m))(*[w for s in suffixes for w in (s & (2**64 - 1), s >> 64)])
    lens = (C.c_uint64 * m)(*prefixes)
    new_lens = (C.c_uint64 * m)()
    ptrs = (C.c_void_p * m)(*[d.ptr for d in devs])
    l0 = hal.launch_count()
    hal._check(hal._lib.b200_fold_multilinears_high_to_low(hal._ctx, ptrs, m, n_vars, lens, sfx, zs, new_lens))
    assert hal.launch_count() - l0 == 1
    for t in range(m):
        exp = oracle.fold_left_lerp_inplace(host[t], prefixes[t], suffixes[t], n_vars, z)
        assert int(new_lens[t]) == len(exp)
        assert _same(hal.to_host(devs[t].slice(0, len(exp))), exp), t
