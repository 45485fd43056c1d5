"""GPU parity tests of the ComputeLayer ops: CUDA path (through the C ABI) vs the CPU oracle on the
same seeded inputs.  Mirrors crates/compute_test_utils/src/layer.rs (instantiated for CpuLayer in
crates/compute/tests/layer.rs:12-165).  Bit-exact (integer GF(2^k) arithmetic)."""
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


def test_copy_fill_roundtrip(hal, oracle):
    x = oracle.rand_b128(1, 1000)
    d = hal.to_device(x)
    assert _same(hal.to_host(d), x)
    d2 = hal.dev_alloc(1000)
    hal.copy_d2d(d, d2)
    assert _same(hal.to_host(d2), x)
    v = 0x0123456789ABCDEF_FEDCBA9876543210
    hal.fill(d2.slice(10, 500), v)
    got = oracle.to_ints(hal.to_host(d2))
    exp = oracle.to_ints(x)
    exp[10:500] = [v] * 490
    assert got == exp
    import binius_b200

    with pytest.raises(binius_b200.InputValidation):
        hal.copy_d2d(d, d2.slice(0, 999))


@pytest.mark.parametrize("n", [1, 2, 7, 32, 1000, 1 << 15, (1 << 17) + 13])
def test_extrapolate_line(hal, oracle, n):
    # cfg#1 of BASELINE.json is n = 2^15 (2^16 input coefficients); compute_test_utils layer.rs:370-420
    e0, e1 = oracle.rand_b128(10 + n, n), oracle.rand_b128(20 + n, n)
    z = oracle.to_ints(oracle.rand_b128(30 + n, 1))[0]
    d0, d1 = hal.to_device(e0), hal.to_device(e1)
    hal.execute(lambda ex: (ex.extrapolate_line(d0, d1, z), [])[1])
    assert _same(hal.to_host(d0), oracle.extrapolate_line(e0, e1, z))
    assert _same(hal.to_host(d1), e1)


def test_extrapolate_line_edge_values(hal, oracle, kat):
    import binius_b200

    g = kat["generators"]["128"]
    special = [0, 1, (1 << 128) - 1, 0xFF, 0xFFFF, 0xFFFFFFFF, (1 << 64) - 1, g, 1 << 127, 1 << 64]
    e0 = oracle.to_arr(special * 4)
    e1 = oracle.to_arr((special[3:] + special[:3]) * 4)
    for z in [0, 1, g, (1 << 128) - 1, 1 << 127, 0x2]:
        d0, d1 = hal.to_device(e0), hal.to_device(e1)
        hal.execute(lambda ex: (ex.extrapolate_line(d0, d1, z), [])[1])
        assert _same(hal.to_host(d0), oracle.extrapolate_line(e0, e1, z))
    with pytest.raises(binius_b200.InputValidation):
        hal.execute(lambda ex: (ex.extrapolate_line(d0, d1.slice(0, 5), 3), [])[1])


def test_extrapolate_line_unaligned_subslices(hal, oracle):
    # ALIGNMENT = 1: split_half below any packing width (memory.rs:165-195)
    x = oracle.rand_b128(77, 64)
    d = hal.to_device(x)
    z = 0xDEADBEEF00000000CAFEBABE12345678
    cur = d.slice(3, 3 + 32)
    ref = x[3:35].copy()
    while cur.len() > 1:
        lo, hi = cur.split_half_mut()
        hal.execute(lambda ex: (ex.extrapolate_line(lo, hi, z), [])[1])
        h = len(ref) // 2
        ref = oracle.extrapolate_line(ref[:h], ref[h:], z)
        cur = lo
        assert _same(hal.to_host(cur), ref)


@pytest.mark.parametrize("log_n,k", [(0, 0), (0, 1), (0, 5), (2, 3), (0, 11), (3, 10), (0, 14), (12, 2), (5, 12), (0, 12), (0, 13), (12, 0), (12, 1),
                                     (13, 3), (7, 15), (0, 22), (3, 21), (0, 25)])
def test_tensor_expand(hal, oracle, log_n, k):
    # compute_test_utils layer.rs:30-71 (zero-filled tail, compare with tensor_prod_eq_ind)
    import random

    rng = random.Random(log_n * 100 + k)
    coords = [rng.getrandbits(128) for _ in range(k)]
    data = np.zeros((1 << (log_n + k), 2), np.uint64)
    data[: 1 << log_n] = oracle.rand_b128(5, 1 << log_n)
    d = hal.to_device(data)
    hal.execute(lambda ex: (ex.tensor_expand(log_n, coords, d), [])[1])
    assert _same(hal.to_host(d), oracle.tensor_expand(data, log_n, coords))


def test_tensor_expand_plans_agree(hal, oracle):
    """the outer-product plan (k_expand_pair + k_expand_outer) against the doubling chain it replaces, incl. zero and one
    coordinates, at the size bench.py times (k = 22)"""
    import random

    rng = random.Random(2222)
    for log_n, k in [(0, 17), (0, 22), (5, 18), (12, 9)]:
        coords = [rng.choice([rng.getrandbits(128)] * 6 + [0, 1]) for _ in range(k)]
        data = np.zeros((1 << (log_n + k), 2), np.uint64)
        data[: 1 << log_n] = oracle.rand_b128(50 + k, 1 << log_n)
        got = []
        for outer in (1, 0):
            hal.set_tuning("expand_outer", outer)
            try:
                d = hal.to_device(data)
                hal.execute(lambda ex: (ex.tensor_expand(log_n, coords, d), [])[1])
                got.append(hal.to_host(d))
            finally:
                hal.set_tuning("expand_outer", 1)
        assert _same(got[0], got[1]), (log_n, k)


def test_tensor_expand_overwrites_and_validates(hal, oracle):
    # the upper part is an OUTPUT (layer.rs:269-296 definition; tensor_prod_eq_ind.rs:70-72 assigns)
    import binius_b200

    coords = [3, 0x1234 << 64, 7, 9, 11, 13, 15, 17, 19, 21, 23, 25, 27]
    for k in (3, 13):
        data = oracle.rand_b128(9, 1 << k)  # garbage beyond element 0
        clean = np.zeros_like(data)
        clean[0] = data[0]
        d = hal.to_device(data)
        hal.execute(lambda ex: (ex.tensor_expand(0, coords[:k], d), [])[1])
        assert _same(hal.to_host(d), oracle.tensor_expand(clean, 0, coords[:k]))
    with pytest.raises(binius_b200.InputValidation):
        hal.execute(lambda ex: (ex.tensor_expand(1, coords[:3], d), [])[1])


def test_eq_ind_partial_eval(hal, oracle):
    # compute/src/ops.rs:26-50
    import binius_b200

    holder_mem = hal.dev_alloc(1 << 10)
    alloc = binius_b200.BumpAllocator(holder_mem)
    point = [5, 1 << 100, 0xABCDEF, 77]
    out = binius_b200.eq_ind_partial_eval(hal, alloc, point)
    exp = oracle.tensor_expand(oracle.to_arr([1] + [0] * 15), 0, point)
    assert _same(hal.to_host(out), exp)
    assert alloc.remaining() == (1 << 10) - 16


@pytest.mark.parametrize("lvl", [0, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("log_b", [7, 10, 14])
def test_inner_product(hal, oracle, lvl, log_b):
    # compute_test_utils layer.rs:73-120
    import binius_b200

    n_b = 1 << log_b
    n_a = n_b >> (7 - lvl)
    a, b = oracle.rand_b128(40 + lvl, n_a), oracle.rand_b128(50 + log_b, n_b)
    da, db = hal.to_device(a), hal.to_device(b)
    (got,) = hal.execute(lambda ex: [ex.inner_product(binius_b200.SubfieldSlice(da, lvl), db)])
    assert got == oracle.inner_product(a, lvl, b)
    with pytest.raises(binius_b200.InputValidation):
        hal.execute(lambda ex: [ex.inner_product(binius_b200.SubfieldSlice(da, lvl), db.slice(0, n_b - 1))])


@pytest.mark.parametrize("lvl", [0, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("log_q", [0, 2, 5])
def test_fold_left_right(hal, oracle, lvl, log_q):
    # compute_test_utils layer.rs:122-260
    import binius_b200

    n_mat = 1 << 6
    log_evals = 6 + 7 - lvl
    mat, vec = oracle.rand_b128(60 + lvl, n_mat), oracle.rand_b128(70 + log_q, 1 << log_q)
    n_out = 1 << (log_evals - log_q)
    dm, dv = hal.to_device(mat), hal.to_device(vec)
    dl, dr = hal.dev_alloc(n_out), hal.dev_alloc(n_out)
    hal.execute(lambda ex: (ex.fold_left(binius_b200.SubfieldSlice(dm, lvl), dv, dl), ex.fold_right(binius_b200.SubfieldSlice(dm, lvl), dv, dr), [])[2])
    assert _same(hal.to_host(dl), oracle.fold_left(mat, lvl, vec, n_out))
    assert _same(hal.to_host(dr), oracle.fold_right(mat, lvl, vec, n_out))
    with pytest.raises(binius_b200.InputValidation):
        hal.execute(lambda ex: (ex.fold_left(binius_b200.SubfieldSlice(dm, lvl), dv, dl.slice(0, n_out - 1)), [])[1])
    hal.dev_free(dl)
    hal.dev_free(dr)


def _expr_cases():
    c = 0x1234567890ABCDEF1122334455667788
    return [
        ([("var", 0), ("var", 1), ("mul", 0, 1)], 2),  # BivariateProduct
        ([("const", 1), ("var", 0), ("mul", 0, 1), ("var", 1), ("mul", 2, 3)], 2),  # IndexComposition<BivariateProduct>
        ([("var", 0), ("var", 1), ("mul", 0, 1), ("const", c), ("add", 2, 3), ("pow", 4, 3), ("var", 2), ("add", 5, 6)], 3),
        ([("var", 0), ("pow", 0, 0)], 1),
        ([("var", 2), ("var", 0), ("add", 0, 1), ("pow", 2, 5)], 3),
    ]


@pytest.mark.parametrize("case", range(5))
def test_compute_composite_and_kernel_sum(hal, oracle, case):
    # compute_test_utils layer.rs:422-520 (compute_composite), :262-330 (kernel add / sum)
    import binius_b200

    steps, n_vars = _expr_cases()[case]
    n = 1 << 9
    ins = [oracle.rand_b128(80 + j, n) for j in range(n_vars)]
    dins = [hal.to_device(a) for a in ins]
    dout = hal.dev_alloc(n)
    ev = hal.compile_expr(binius_b200.ArithCircuit(steps))
    hal.execute(lambda ex: (ex.compute_composite(binius_b200.SlicesBatch(dins, n), dout, ev), [])[1])
    assert _same(hal.to_host(dout), oracle.compute_composite(ins, steps))
    coeff, init = 0xABCDEF0123456789 << 50, 0x77

    def kern(kex, log_chunks, bufs):
        acc = kex.decl_value(init)
        kex.sum_composition_evals(binius_b200.SlicesBatch([b.to_ref() for b in bufs], n), ev, coeff, acc)
        return [acc]

    (got,) = hal.execute(lambda ex: ex.accumulate_kernels(kern, [binius_b200.KernelMemMap.Chunked(d, 3) for d in dins]))
    assert got == oracle.sum_composition_evals(ins, steps, coeff, init)
    with pytest.raises(binius_b200.InputValidation):
        hal.execute(lambda ex: (ex.compute_composite(binius_b200.SlicesBatch(dins, n), dout.slice(0, n - 1), ev), [])[1])


def test_kernel_add_and_map_kernels(hal, oracle):
    # compute_test_utils layer.rs:262-330: map_kernels with Chunked, ChunkedMut and Local buffers
    import binius_b200

    n, log_n = 1 << 10, 10
    a, b = oracle.rand_b128(90, n), oracle.rand_b128(91, n)
    da, db, dc = hal.to_device(a), hal.to_device(b), hal.dev_alloc(n)

    def kern(kex, log_chunks, bufs):
        la = log_n - log_chunks
        kex.add(la, bufs[0].to_ref(), bufs[1].to_ref(), bufs[3].data)  # local = a + b
        kex.add_assign(la, bufs[0].to_ref(), bufs[3].data)  # local += a  -> b
        kex.add(la, bufs[3].to_ref(), bufs[0].to_ref(), bufs[2].data)  # c = b + a
        return None

    M = binius_b200.KernelMemMap
    hal.execute(lambda ex: (ex.map_kernels(kern, [M.Chunked(da, 0), M.Chunked(db, 0), M.ChunkedMut(dc, 0), M.Local(log_n)]), [])[1])
    assert _same(hal.to_host(dc), a ^ b)
    assert M.log_chunks_range([M.Chunked(da, 3), M.Local(6)]) == (0, 6)
    assert M.log_chunks_range([M.Chunked(da, 3), M.Local(9)]) == (0, 7)


@pytest.mark.parametrize("log_n", [1, 4, 11])
def test_pairwise_product_reduce(hal, oracle, log_n):
    # compute_test_utils layer.rs:522-600
    import binius_b200

    x = oracle.rand_b128(100 + log_n, 1 << log_n)
    dx = hal.to_device(x)
    outs = [hal.dev_alloc(1 << (log_n - r - 1)) for r in range(log_n)]
    hal.execute(lambda ex: (ex.pairwise_product_reduce(dx, outs), [])[1])
    for got, exp in zip(outs, oracle.pairwise_product_reduce(x)):
        assert _same(hal.to_host(got), exp)
    with pytest.raises(binius_b200.InputValidation):
        hal.execute(lambda ex: (ex.pairwise_product_reduce(dx, outs[:-1] if log_n > 1 else outs + outs), [])[1])
    with pytest.raises(binius_b200.InputValidation):
        hal.execute(lambda ex: (ex.pairwise_product_reduce(dx.slice(0, 1), []), [])[1])


@pytest.mark.parametrize("n_vars,m", [(1, 2), (6, 4), (12, 8), (13, 3), (15, 5)])
def test_bivariate_round_evals_fused_and_traced(hal, oracle, n_vars, m):
    # core/src/protocols/sumcheck/v3/bivariate_product.rs:303-408, both as the fused entry point and
    # as the literal accumulate_kernels program the reference prover issues
    import random

    import binius_b200

    rng = random.Random(n_vars)
    mls = [oracle.rand_b128(110 + t, 1 << n_vars) for t in range(m)]
    dmls = [hal.to_device(x) for x in mls]
    pairs = [(rng.randrange(m), rng.randrange(m)) for _ in range(m)]
    alpha = rng.getrandbits(128)
    exp = oracle.bivariate_round_evals(mls, n_vars, pairs, alpha)
    got = hal.execute(lambda ex: list(ex.bivariate_round_evals(dmls, n_vars, pairs, alpha)))
    assert got == exp

    half_log = n_vars - 1
    M, SB = binius_b200.KernelMemMap, binius_b200.SlicesBatch
    ev = hal.compile_expr(binius_b200.ArithCircuit([("var", 0), ("var", 1), ("mul", 0, 1)]))
    pows = [1]
    for _ in pairs:
        pows.append(oracle.mul(pows[-1], alpha))

    def kern(kex, log_chunks, bufs):
        sz = half_log - log_chunks
        lo = [bufs[3 * t] for t in range(m)]
        hi = [bufs[3 * t + 1] for t in range(m)]
        inf = [bufs[3 * t + 2] for t in range(m)]
        y1 = kex.decl_value(0)
        for c, (ia, ib) in enumerate(pairs):
            kex.sum_composition_evals(SB([hi[ia].to_ref(), hi[ib].to_ref()], 1 << sz), ev, pows[c], y1)
        for t in range(m):
            kex.add(sz, lo[t].to_ref(), hi[t].to_ref(), inf[t].data)
        yinf = kex.decl_value(0)
        for c, (ia, ib) in enumerate(pairs):
            kex.sum_composition_evals(SB([inf[ia].to_ref(), inf[ib].to_ref()], 1 << sz), ev, pows[c], yinf)
        return [y1, yinf]

    maps = []
    for d in dmls:
        lo, hi = d.split_half()
        maps += [M.Chunked(lo, 0), M.Chunked(hi, 0), M.Local(half_log)]
    got2 = hal.execute(lambda ex: ex.accumulate_kernels(kern, maps))
    assert got2 == exp


def test_fri_fold(hal, oracle):
    # compute_test_utils layer.rs:332-368 ; cpu/layer.rs:304-391
    import random

    import binius_b200

    ntt = binius_b200.B200AdditiveNTT(hal, 5, 12)
    ontt = oracle.NTT(5, 12)
    rng = random.Random(5)
    for log_len, log_batch, n_ch in [(6, 2, 2), (6, 0, 3), (5, 2, 5), (4, 3, 3), (3, 0, 0), (10, 4, 4), (12, 0, 4),
                                     (12, 1, 1), (12, 2, 2), (13, 3, 3), (12, 4, 4)]:  # last four: K64 tensor-lerp kernel
        ch = [rng.getrandbits(128) for _ in range(n_ch)]
        data = oracle.rand_b128(log_len * 10 + n_ch, 1 << (log_len + log_batch))
        n_out = 1 << (log_len - (n_ch - log_batch))
        din, dout = hal.to_device(data), hal.dev_alloc(n_out)
        hal.execute(lambda ex: (ex.fri_fold(ntt, log_len, log_batch, ch, din, dout), [])[1])
        assert _same(hal.to_host(dout), ontt.fri_fold(log_len, log_batch, ch, data, n_out))
    with pytest.raises(binius_b200.InputValidation):
        hal.execute(lambda ex: (ex.fri_fold(ntt, 4, 2, [1], din.slice(0, 64), dout.slice(0, 16)), [])[1])
    for i in range(1, 13):
        for j in (0, 1, 5, 100):
            assert ntt.get_subspace_eval(i, j) == ontt.get_subspace_eval(i, j)


def test_holder_and_bump_allocator(oracle):
    # ComputeHolder::to_data + BumpAllocator OutOfMemory (alloc.rs:110-114)
    import binius_b200

    holder = binius_b200.B200LayerHolder.new(1 << 8, 1 << 12)
    data = holder.to_data()
    a = data.dev_alloc.alloc(1 << 11)
    b = data.dev_alloc.alloc(1 << 11)
    assert a.ptr + 16 * (1 << 11) == b.ptr
    with pytest.raises(binius_b200.AllocError):
        data.dev_alloc.alloc(1)
    h = data.host_alloc.alloc(16)
    h[:] = oracle.rand_b128(3, 16)
    data.hal.copy_h2d(h, a.slice(0, 16))
    assert _same(data.hal.to_host(a.slice(0, 16)), h)
    assert _same(data.hal.to_host(a.slice(16, 32)), np.zeros((16, 2), np.uint64))
    holder.layer.close()


@pytest.mark.parametrize("n", [1, 1000, (1 << 20) + 5, 3 << 20])
def test_extrapolate_line_host_pipeline(hal, oracle, n):
    # host-buffer form (old-HAL shaped): chunked H2D/kernel/D2H pipeline, pageable and pinned memory
    e0, e1 = oracle.rand_b128(700 + n % 97, n), oracle.rand_b128(701 + n % 97, n)
    z = 0x0123456789ABCDEF_0FEDCBA987654321
    sample = slice(0, min(n, 1 << 12))
    exp_head = oracle.extrapolate_line(e0[sample], e1[sample], z)
    exp_tail = oracle.extrapolate_line(e0[-257:], e1[-257:], z)
    a = e0.copy()
    hal.extrapolate_line_host(a, e1, z)
    assert _same(a[sample], exp_head) and _same(a[-257:], exp_tail)
    p0, p1 = hal.host_alloc(n), hal.host_alloc(n)
    p0[:], p1[:] = e0, e1
    hal.extrapolate_line_host(p0, p1, z)
    assert np.array_equal(p0, a)
    d0, d1 = hal.to_device(e0), hal.to_device(e1)
    hal.execute(lambda ex: (ex.extrapolate_line(d0, d1, z), [])[1])
    assert _same(hal.to_host(d0), a)
    hal.dev_free(d0)
    hal.dev_free(d1)


def test_deferred_fold_batching(hal, oracle):
    """consecutive extrapolate_line calls with one challenge are deferred and launched as ONE
    multi-segment kernel; dependent folds, a different challenge or any other op flush the queue
    (store-to-load order of compute/src/layer.rs:90-99 must be preserved)"""
    m, n = 60, 1 << 9
    z1, z2 = 0x1111111122222222_3333333344444444, 0x5555555566666666_7777777788888889
    hosts = [oracle.rand_b128(800 + t, n) for t in range(m)]
    devs = [hal.to_device(h) for h in hosts]
    l0 = hal.launch_count()

    def op(ex):
        for d in devs:
            lo, hi = d.split_half_mut()
            ex.extrapolate_line(lo, hi, z1)
        return []

    hal.execute(op)
    assert hal.launch_count() - l0 <= 2  # 60 folds -> two launches of <= 48 segments
    exp = [oracle.extrapolate_line(h[: n // 2], h[n // 2:], z1) for h in hosts]
    for d, e in zip(devs, exp):
        assert _same(hal.to_host(d.slice(0, n // 2)), e)

    # dependent chain + challenge switch + interleaved copy inside one execute
    d, h = devs[0], exp[0]
    scratch = hal.dev_alloc(n // 4)

    def op2(ex):
        a, b = d.slice(0, n // 4), d.slice(n // 4, n // 2)
        ex.extrapolate_line(a, b, z1)  # queued
        a2, b2 = d.slice(0, n // 8), d.slice(n // 8, n // 4)
        ex.extrapolate_line(a2, b2, z1)  # reads the pending output -> flush first
        ex.extrapolate_line(d.slice(n // 4, n // 4 + 8), d.slice(n // 2, n // 2 + 8), z2)  # new challenge -> flush
        hal.copy_d2d(d.slice(0, n // 4), scratch)  # any other op -> flush
        return []

    hal.execute(op2)
    r1 = oracle.extrapolate_line(h[: n // 4], h[n // 4: n // 2], z1)
    r2 = oracle.extrapolate_line(r1[: n // 8], r1[n // 8: n // 4], z1)
    r3 = oracle.extrapolate_line(h[n // 4: n // 4 + 8], hosts[0][n // 2: n // 2 + 8], z2)
    got = hal.to_host(d.slice(0, n // 2))
    assert _same(got[: n // 8], r2) and _same(got[n // 8: n // 4], r1[n // 8:]) and _same(got[n // 4: n // 4 + 8], r3)
    assert _same(hal.to_host(scratch)[: n // 8], r2)


@pytest.mark.parametrize("lvl", [0, 3, 4, 5, 6, 7])
def test_fold_right_ring_switch_shape(hal, oracle, lvl):
    # ring-switch shape (core/src/ring_switch/eq_ind.rs:140-146): vec.len() * 2^lvl == 128 -> LUT fast path
    import binius_b200

    n = 1 << 12
    mat, vec = oracle.rand_b128(910 + lvl, n), oracle.rand_b128(920 + lvl, 128 >> lvl)
    dm, dv, do = hal.to_device(mat), hal.to_device(vec), hal.dev_alloc(n)
    hal.execute(lambda ex: (ex.fold_right(binius_b200.SubfieldSlice(dm, lvl), dv, do), [])[1])
    assert _same(hal.to_host(do), oracle.fold_right(mat, lvl, vec, n))


def _xor_sum(arr) -> int:
    a = np.asarray(arr, dtype=np.uint64).reshape(-1, 2)
    lo, hi = np.bitwise_xor.reduce(a[:, 0]), np.bitwise_xor.reduce(a[:, 1])
    return int(lo) | (int(hi) << 64)


def test_fold_high_full_size_properties(hal, oracle):
    """BASELINE metric size (2^24 input coefficients) through size-independent properties: slabs against
    the oracle, and the XOR checksum  sum(out) = sum(e0) + z * (sum(e0) + sum(e1))  (multiplication by the
    challenge is GF(2)-linear), plus the degenerate challenges 0 and 1."""
    n = 1 << 23
    rng = np.random.default_rng(2024)
    e0 = rng.integers(0, 1 << 63, size=(n, 2), dtype=np.int64).astype(np.uint64) * np.uint64(2) + np.uint64(1)
    e1 = rng.integers(0, 1 << 63, size=(n, 2), dtype=np.int64).astype(np.uint64) * np.uint64(3)
    z = 0x2E895399AF449ACE499596F6E5FCCAFA
    d0, d1 = hal.to_device(e0), hal.to_device(e1)
    hal.execute(lambda ex: (ex.extrapolate_line(d0, d1, z), [])[1])
    out = hal.to_host(d0)
    for off in (0, 4096 * 777 + 3, n - 5000):
        sl = slice(off, off + 4321)
        assert _same(out[sl], oracle.extrapolate_line(e0[sl], e1[sl], z))
    s0, s1 = _xor_sum(e0), _xor_sum(e1)
    assert _xor_sum(out) == s0 ^ oracle.mul(s0 ^ s1, z)
    assert _same(hal.to_host(d1), e1)
    # z = 1 selects e1, z = 0 keeps e0
    hal.execute(lambda ex: (ex.extrapolate_line(d0, d1, 1), [])[1])
    assert _same(hal.to_host(d0), e1)
    hal.copy_h2d(e0, d0)
    hal.execute(lambda ex: (ex.extrapolate_line(d0, d1, 0), [])[1])
    assert _same(hal.to_host(d0), e0)


def test_fold_multilinears_large_truncated_both_orders(hal, oracle):
    """2^20-variable-count-scale multilinears with ragged stored prefixes and constant suffixes through
    the TMA fold kernels (tiles that end mid-warp, pivot inside a tile, empty second operand), both
    evaluation orders, whole result against the oracle."""
    import ctypes as C

    n_vars = 18
    full = 1 << n_vars
    prefixes = [full, full - 1, full // 2 + 77, full // 2, full // 2 - 1, 64 * 1000 + 33, 65, 1]
    suffixes = [0, 5, 1 << 100, 7, 0x1234, (1 << 128) - 1, 3, 9]
    z = 0x0F1E2D3C4B5A69788796A5B4C3D2E1F1
    host = [oracle.rand_b128(4000 + t, p) for t, p in enumerate(prefixes)]
    m = len(host)
    zs = (C.c_uint64 * 2)(z & (2**64 - 1), z >> 64)
    sfx = (C.c_uint64 * (2 * m))(*[w for s in suffixes for w in (s & (2**64 - 1), s >> 64)])
    lens = (C.c_uint64 * m)(*prefixes)
    new_lens = (C.c_uint64 * m)()
    # HighToLow, in place
    dev = [hal.to_device(h) for h in host]
    ptrs = (C.c_void_p * m)(*[d.ptr for d in dev])
    hal._check(hal._lib.b200_fold_multilinears_high_to_low(hal._ctx, ptrs, m, n_vars, lens, sfx, zs, new_lens))
    for t in range(m):
        exp = oracle.fold_left_lerp_inplace(host[t], prefixes[t], suffixes[t], n_vars, z)
        assert int(new_lens[t]) == len(exp)
        assert _same(hal.to_host(dev[t].slice(0, len(exp))), exp), t
    # LowToHigh, out of place
    dev = [hal.to_device(h) for h in host]
    outs = [hal.dev_alloc((p + 1) // 2) for p in prefixes]
    ptrs = (C.c_void_p * m)(*[d.ptr for d in dev])
    optrs = (C.c_void_p * m)(*[o.ptr for o in outs])
    hal._check(hal._lib.b200_fold_multilinears_low_to_high(hal._ctx, ptrs, optrs, m, n_vars, lens, sfx, zs, new_lens))
    for t in range(m):
        exp = oracle.fold_right_lerp(host[t], suffixes[t], z)
        assert int(new_lens[t]) == len(exp)
        assert _same(hal.to_host(outs[t].slice(0, len(exp))), exp), t
        assert _same(hal.to_host(dev[t]), host[t])  # inputs untouched


@pytest.mark.parametrize("log_q", [0, 3, 7])
def test_fold_left_b1_fast_path(hal, oracle, log_q):
    """evaluate_partial_high of a bit-packed (B1) multilinear by a short tensor query: the
    warp-transposed byte-LUT kernel (n_out >= 4096, a multiple of 128) against the oracle's fold_left."""
    import binius_b200

    log_out = 13
    n_mat = 1 << (log_out + log_q - 7)
    mat, vec = oracle.rand_b128(800 + log_q, n_mat), oracle.rand_b128(810 + log_q, 1 << log_q)
    dm, dv, do = hal.to_device(mat), hal.to_device(vec), hal.dev_alloc(1 << log_out)
    hal.execute(lambda ex: (ex.fold_left(binius_b200.SubfieldSlice(dm, 0), dv, do), [])[1])
    assert _same(hal.to_host(do), oracle.fold_left(mat, 0, vec, 1 << log_out))


def test_inner_product_b128_large(hal, oracle):
    """B128 x B128 inner products of >= 4096 elements run as tensor-core jobs (split in two halves)."""
    import binius_b200

    for n in (4096, 4096 + 64, 1 << 15):
        a, b = oracle.rand_b128(820 + n % 7, n), oracle.rand_b128(830 + n % 5, n)
        da, db = hal.to_device(a), hal.to_device(b)
        got = hal.execute(lambda ex: [ex.inner_product(binius_b200.SubfieldSlice(da, 7), db)])
        assert got == [oracle.inner_product(a, 7, b)]


@pytest.mark.parametrize("n_vars", [3, 9, 13])
def test_kernel_scope_general_programs(hal, oracle, n_vars):
    """accumulate_kernels programs other than the prover's bivariate closure (layer.rs:134-245): degree-1 and constant
    monomials, a scaled monomial, a square, add_assign into a Local, a sum over a Local before and after it is written
    (fused lowering), and a program that writes a ChunkedMut buffer (runs op by op).  Checked against the oracle's
    sum_composition_evals on host-side copies of the buffers."""
    import random

    import binius_b200
    from binius_b200 import ArithCircuit as A

    rng = random.Random(900 + n_vars)
    n = 1 << n_vars
    m = 3
    host = [oracle.rand_b128(700 + t, n) for t in range(m)]
    dev = [hal.to_device(x) for x in host]
    M, SB = binius_b200.KernelMemMap, binius_b200.SlicesBatch
    exprs = [A.var(0) * A.var(1) + A.var(2) + A.constant(rng.getrandbits(128)),
             A.constant(rng.getrandbits(128)) * A.var(2) * A.var(2) + A.var(0),
             A.var(1)]
    evs = [hal.compile_expr(e) for e in exprs]
    coeffs = [rng.getrandbits(128) for _ in exprs]
    init = rng.getrandbits(128)
    xor = lambda a, b: a ^ b  # noqa: E731

    # (1) fusable: sums over mapped buffers, add into Locals, sums over the Locals
    def kern(kex, log_chunks, bufs):
        a, b, c, l0, l1, l2 = bufs
        acc = kex.decl_value(init)
        for ev, cf in zip(evs, coeffs):
            kex.sum_composition_evals(SB([a.to_ref(), b.to_ref(), c.to_ref()], n), ev, cf, acc)
        kex.add(n_vars, a.to_ref(), b.to_ref(), l0.data)
        kex.add(n_vars, b.to_ref(), c.to_ref(), l1.data)
        kex.add(n_vars, c.to_ref(), a.to_ref(), l2.data)
        acc2 = kex.decl_value(0)
        for ev, cf in zip(evs, coeffs):
            kex.sum_composition_evals(SB([l0.to_ref(), l1.to_ref(), c.to_ref()], n), ev, cf, acc2)
        kex.sum_composition_evals(SB([l2.to_ref(), l2.to_ref(), l0.to_ref()], n), evs[0], coeffs[1], acc2)
        return [acc, acc2]

    maps = [M.Chunked(d, 0) for d in dev] + [M.Local(n_vars)] * 3
    got = hal.execute(lambda ex: ex.accumulate_kernels(kern, maps))
    l0, l1, l2 = host[0] ^ host[1], host[1] ^ host[2], host[2] ^ host[0]
    exp1 = init
    for e, cf in zip(exprs, coeffs):
        exp1 = oracle.sum_composition_evals(host, e.steps, cf, exp1)
    exp2 = 0
    for e, cf in zip(exprs, coeffs):
        exp2 = oracle.sum_composition_evals([l0, l1, host[2]], e.steps, cf, exp2)
    exp2 = oracle.sum_composition_evals([l2, l2, l0], exprs[0].steps, coeffs[1], exp2)
    assert got == [exp1, exp2]

    # (2) not fusable: add_assign into a Local that is read before (zero) and after, and a write into a ChunkedMut buffer
    out = hal.dev_alloc(n)
    hal.fill(out, 0)

    def kern2(kex, log_chunks, bufs):
        a, b, c, o, l0 = bufs
        acc = kex.decl_value(0)
        kex.sum_composition_evals(SB([l0.to_ref(), a.to_ref(), b.to_ref()], n), evs[0], coeffs[0], acc)  # Local still zero
        kex.add_assign(n_vars, a.to_ref(), l0.data)
        kex.add_assign(n_vars, b.to_ref(), l0.data)
        kex.sum_composition_evals(SB([l0.to_ref(), a.to_ref(), b.to_ref()], n), evs[0], coeffs[2], acc)
        kex.add(n_vars, l0.to_ref(), c.to_ref(), o.data)
        return [acc]

    maps2 = [M.Chunked(d, 0) for d in dev] + [M.ChunkedMut(out, 0), M.Local(n_vars)]
    got2 = hal.execute(lambda ex: ex.accumulate_kernels(kern2, maps2))
    zero = np.zeros_like(host[0])
    e = oracle.sum_composition_evals([zero, host[0], host[1]], exprs[0].steps, coeffs[0], 0)
    e = oracle.sum_composition_evals([l0, host[0], host[1]], exprs[0].steps, coeffs[2], e)
    assert got2 == [e]
    assert np.array_equal(hal.to_host(out), l0 ^ host[2])


def test_one_layer_from_several_host_threads(hal, oracle):
    """ComputeLayer is used as `&self` from several host threads (rayon join / map, layer.rs:115-131): the context
    serialises its entry points and an `execute` is a scope over the result slots, so concurrent executes on ONE layer
    take turns and every thread gets its own values (ctypes releases the GIL inside the library calls)."""
    import threading

    import binius_b200

    n = 1 << 14
    jobs = []
    for t in range(4):
        a, b = oracle.rand_b128(9100 + t, n), oracle.rand_b128(9200 + t, n)
        z = 0x1F2E3D4C5B6A7988 + t
        jobs.append((a, b, z, hal.to_device(a), hal.to_device(b), hal.to_device(a), hal.to_device(b)))
    errors = []

    def work(t):
        try:
            a, b, z, da, db, fa, fb = jobs[t]
            for it in range(10):
                got = hal.execute(lambda ex: [ex.inner_product(binius_b200.SubfieldSlice(da, 7), db)])
                assert got == [oracle.inner_product(a, 7, b)], (t, it)
            hal.execute(lambda ex: (ex.extrapolate_line(fa, fb, z), [])[1])
            assert _same(hal.to_host(fa), oracle.extrapolate_line(a, b, z)), t
        except Exception as e:  # surfaced in the main thread
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_fold_right_same_query_batches_and_keeps_order(hal, oracle):
    """fold_right by a 128-coefficient query over B1 matrices is deferred: calls sharing the query go out as one launch
    (launch_count), a call that reads an earlier output or uses another query flushes first; values vs the oracle."""
    import binius_b200

    S = binius_b200.SubfieldSlice
    n = 1 << 11
    q1, q2 = oracle.rand_b128(9001, 128), oracle.rand_b128(9002, 128)
    mats = [oracle.rand_b128(9100 + j, n) for j in range(5)]
    dq1, dq2 = hal.to_device(q1), hal.to_device(q2)
    dm = [hal.to_device(m) for m in mats]
    outs = [hal.dev_alloc(n) for _ in range(6)]
    hal.sync()
    before = hal.launch_count()

    def op(ex):
        for j in range(4):
            ex.fold_right(S(dm[j], 0), dq1, outs[j])        # one batch of four
        ex.fold_right(S(outs[0], 0), dq1, outs[4])          # reads an output of the batch: flush, then a new batch
        ex.fold_right(S(dm[4], 0), dq2, outs[5])            # another query: flush
        return []

    hal.execute(op)
    hal.sync()
    assert hal.launch_count() - before == 3
    for j in range(4):
        assert _same(hal.to_host(outs[j]), oracle.fold_right(mats[j], 0, q1, n))
    assert _same(hal.to_host(outs[4]), oracle.fold_right(oracle.fold_right(mats[0], 0, q1, n), 0, q1, n))
    assert _same(hal.to_host(outs[5]), oracle.fold_right(mats[4], 0, q2, n))
