"""Oracle op restatements vs independent pure-Python definitions and the algebraic properties the
reference's own tests assert.  CPU only, small sizes."""
import random

import numpy as np
import pytest


def ints(o, a):
    return o.to_ints(a)


def test_extrapolate_line_definition(oracle):
    # cpu/layer.rs:393-408
    o = oracle
    e0, e1 = o.rand_b128(0, 33), o.rand_b128(1, 33)
    z = ints(o, o.rand_b128(2, 1))[0]
    got = ints(o, o.extrapolate_line(e0, e1, z))
    for g, a, b in zip(got, ints(o, e0), ints(o, e1)):
        assert g == a ^ o.mul(a ^ b, z)
    with pytest.raises(o.OracleError):
        o.extrapolate_line(e0, e1[:-1], z)


def test_tensor_expand_symbolic(oracle):
    # math/src/tensor_prod_eq_ind.rs:113-186 : eq_ind(r)[i] = prod_k (r_k if bit k of i else 1-r_k)
    o = oracle
    rng = random.Random(0)
    for k in range(0, 5):
        r = [rng.getrandbits(128) for _ in range(k)]
        data = np.zeros((1 << k, 2), np.uint64)
        data[0, 0] = 1
        got = ints(o, o.tensor_expand(data, 0, r))
        for i in range(1 << k):
            e = 1
            for b in range(k):
                e = o.mul(e, r[b] if (i >> b) & 1 else r[b] ^ 1)
            assert got[i] == e
        assert k == 0 or len(set(got)) > 1
    # log_n > 0 prefix is scaled element-wise; wrong length is rejected (cpu/layer.rs:288-290)
    v = o.rand_b128(5, 4)
    data = np.zeros((16, 2), np.uint64)
    data[:4] = v
    r = [rng.getrandbits(128) for _ in range(2)]
    got = ints(o, o.tensor_expand(data, 2, r))
    eq = ints(o, o.tensor_expand(o.to_arr([1, 0, 0, 0]), 0, r))
    for hi in range(4):
        for lo in range(4):
            assert got[hi * 4 + lo] == o.mul(ints(o, v)[lo], eq[hi])
    with pytest.raises(o.OracleError):
        o.tensor_expand(data[:8], 2, r)


@pytest.mark.parametrize("lvl", [0, 3, 4, 5, 6, 7])
def test_inner_product_and_folds(oracle, lvl):
    # cpu/layer.rs:205-236, 574-675
    o = oracle
    L = 1 << (7 - lvl)
    w = 1 << lvl
    n_a = 8
    a = o.rand_b128(10 + lvl, n_a)
    b = o.rand_b128(20 + lvl, n_a * L)
    ai, bi = ints(o, a), ints(o, b)
    limbs = [(x >> (j * w)) & ((1 << w) - 1) for x in ai for j in range(L)]
    exp = 0
    for l, y in zip(limbs, bi):
        exp ^= o.mul(y, l)
    assert o.inner_product(a, lvl, b) == exp
    with pytest.raises(o.OracleError):
        o.inner_product(a, lvl, b[:-1])
    # fold_left / fold_right over the flattened limb matrix
    n_evals = n_a * L
    for log_q in range(0, 4):
        q = o.rand_b128(30 + log_q, 1 << log_q)
        qi = ints(o, q)
        rows = n_evals >> log_q
        if rows == 0:
            continue
        fl = ints(o, o.fold_left(a, lvl, q, rows))
        fr = ints(o, o.fold_right(a, lvl, q, rows))
        for i in range(rows):
            el, er = 0, 0
            for j in range(1 << log_q):
                el ^= o.mul(qi[j], limbs[j * rows + i])
                er ^= o.mul(qi[j], limbs[i * (1 << log_q) + j])
            assert fl[i] == el and fr[i] == er
        with pytest.raises(o.OracleError):
            o.fold_left(a, lvl, q, rows + 1)


def test_composite_and_kernel_ops(oracle):
    o = oracle
    n = 16
    ins = [o.rand_b128(40 + j, n) for j in range(3)]
    c = 0x1234567890ABCDEF1122334455667788
    # (x0 * x1 + c) ^ 3 + x2
    steps = [("var", 0), ("var", 1), ("mul", 0, 1), ("const", c), ("add", 2, 3), ("pow", 4, 3), ("var", 2), ("add", 5, 6)]
    got = ints(o, o.compute_composite(ins, steps))
    vals = [ints(o, a) for a in ins]
    for i in range(n):
        t = o.mul(vals[0][i], vals[1][i]) ^ c
        t = o.mul(o.mul(t, t), t)
        assert got[i] == t ^ vals[2][i]
    coeff, acc0 = 0xABCDEF, 0x77
    s = 0
    for g in got:
        s ^= g
    assert o.sum_composition_evals(ins, steps, coeff, acc0) == acc0 ^ o.mul(s, coeff)
    with pytest.raises(o.OracleError):
        o.compute_composite(ins, steps, n_out=n - 1)


def test_pairwise_product_reduce(oracle):
    o = oracle
    x = o.rand_b128(50, 16)
    outs = o.pairwise_product_reduce(x)
    assert [len(t) for t in outs] == [8, 4, 2, 1]
    cur = ints(o, x)
    for t in outs:
        cur = [o.mul(cur[2 * i], cur[2 * i + 1]) for i in range(len(cur) // 2)]
        assert ints(o, t) == cur
    with pytest.raises(o.OracleError):
        o.pairwise_product_reduce(x[:12])
    with pytest.raises(o.OracleError):
        o.pairwise_product_reduce(x[:1])
    with pytest.raises(o.OracleError):
        o.pairwise_product_reduce(x, out_lens=[8, 4, 2])


def test_bivariate_round_evals(oracle):
    # v3/bivariate_product.rs:303-424 : y1 + y0 = claimed sum  (y0 = sum lo_a*lo_b)
    o = oracle
    n_vars, m = 5, 4
    mls = [o.rand_b128(60 + t, 1 << n_vars) for t in range(m)]
    pairs = [(0, 1), (2, 3), (1, 1), (3, 0)]
    alpha = 0xDEADBEEFCAFEBABE0123456789ABCDEF
    y1, yinf = o.bivariate_round_evals(mls, n_vars, pairs, alpha)
    half = 1 << (n_vars - 1)
    e1 = einf = 0
    pw = 1
    for ia, ib in pairs:
        a, b = ints(o, mls[ia]), ints(o, mls[ib])
        s1 = sinf = 0
        for i in range(half):
            s1 ^= o.mul(a[half + i], b[half + i])
            sinf ^= o.mul(a[i] ^ a[half + i], b[i] ^ b[half + i])
        e1 ^= o.mul(s1, pw)
        einf ^= o.mul(sinf, pw)
        pw = o.mul(pw, alpha)
    assert (y1, yinf) == (e1, einf)


# ---------------------------------------------------------------------------------------------- NTT
@pytest.mark.parametrize("kt,d", [(3, 8), (4, 10), (5, 12)])
def test_ntt_twiddle_properties(oracle, kt, d):
    # twiddle.rs:388-534 : What_i vanishes on U_i, is 1 on beta_i, and is GF(2)-linear
    o = oracle
    ntt = o.NTT(kt, d)
    s = ntt.s_evals()
    assert [len(r) for r in s] == [d - 1 - r for r in range(d)]
    assert s[0] == [1 << (j + 1) for j in range(d - 1)]
    # recompute What_r(beta_j) from the product definition for small r
    for r in range(0, min(d - 1, 5)):
        U = [0]
        for b in range(r):
            U = U + [u ^ (1 << b) for u in U]

        def W(x):
            p = 1
            for u in U:
                p = o.mul(p, x ^ u, kt)
            return p

        norm = o.invert(W(1 << r), kt)
        for j in range(r + 1, d):
            assert s[r][j - r - 1] == o.mul(W(1 << j), norm, kt)
    # get_subspace_eval(i, j) = s_evals[d-i].get(j)
    for i in range(1, d + 1):
        row = s[d - i]
        for j in [0, 1, 2, 3, 5]:
            if j >> len(row):
                continue
            e = 0
            for b in range(len(row)):
                if (j >> b) & 1:
                    e ^= row[b]
            assert ntt.get_subspace_eval(i, j) == e


def test_ntt_is_polynomial_evaluation(oracle):
    """Forward NTT output[y] = f(point y of the coset) where f has novel-basis coefficients = input
    (ntt/src/tests/ntt_tests.rs checks against SimpleAdditiveNTT; here against the definition)."""
    o = oracle
    kt, d, log_n = 4, 8, 4
    ntt = o.NTT(kt, d)
    rng = random.Random(7)
    coeffs = np.array([rng.getrandbits(16) for _ in range(1 << log_n)], dtype=np.uint16)
    for coset_bits, coset in [(0, 0), (2, 3), (4, 9)]:
        out = ntt.forward(coeffs, kd=4, log_y=log_n, coset=coset, coset_bits=coset_bits)
        # normalised subspace polys What_b for the full domain rows 0..log_n-1 require row0 = 0,
        # i.e. log_n + coset_bits == d; otherwise the transform uses rows shifted by row0 and the
        # evaluation domain is the corresponding quotient -- covered by the round-trip test below.
        if log_n + coset_bits != d:
            continue

        def What(b, x):
            U = [0]
            for t in range(b):
                U = U + [u ^ (1 << t) for u in U]
            p, q = 1, 1
            for u in U:
                p = o.mul(p, x ^ u, kt)
                q = o.mul(q, (1 << b) ^ u, kt)
            return o.mul(p, o.invert(q, kt), kt)

        for y in range(1 << log_n):
            x = (coset << log_n) | y
            acc = 0
            for c in range(1 << log_n):
                term = int(coeffs[c])
                for b in range(log_n):
                    if (c >> b) & 1:
                        term = o.mul(term, What(b, x), kt)
                acc ^= term
            assert int(out[y]) == acc


@pytest.mark.parametrize("kt,kd,dt", [(3, 3, np.uint8), (4, 4, np.uint16), (5, 5, np.uint32), (5, 7, None), (3, 5, np.uint32)])
def test_ntt_roundtrip_shapes_linearity(oracle, kt, kd, dt):
    # ntt_tests.rs:75-190 (all shapes / cosets / skip_rounds round-trip), single_threaded.rs:498-540
    o = oracle
    d = 8 if kt == 3 else 10
    ntt = o.NTT(kt, d)
    rng = np.random.default_rng(0)
    for (lx, ly, lz, cb, skip) in [(0, 5, 0, 0, 0), (2, 4, 1, 2, 0), (3, 3, 0, 1, 1), (0, 6, 2, 0, 2), (1, 5, 0, 3, 5)]:
        n = 1 << (lx + ly + lz)
        if kd == 7:
            data = o.rand_b128(lx * 100 + ly, n)
        else:
            data = rng.integers(0, 1 << (1 << kd), size=n, dtype=np.uint64).astype(dt)
        coset = (1 << cb) - 1
        f = ntt.forward(data, kd, lx, ly, lz, coset, cb, skip)
        b = ntt.inverse(f, kd, lx, ly, lz, coset, cb, skip)
        assert np.array_equal(b, data)
        if skip < ly:
            assert not np.array_equal(f, data)
        # linearity over GF(2)
        data2 = np.roll(data, 3, axis=0)
        f2 = ntt.forward(data2, kd, lx, ly, lz, coset, cb, skip)
        f12 = ntt.forward(data ^ data2, kd, lx, ly, lz, coset, cb, skip)
        assert np.array_equal(f12, f ^ f2)
    # error classes (single_threaded.rs:364-406)
    data = np.zeros(64, dtype=np.uint32) if kd != 7 else np.zeros((64, 2), np.uint64)
    if kd in (5, 7) and kt == 5:
        with pytest.raises(o.OracleError) as e:
            ntt.forward(data, kd, 0, 6, 0, 0, 0, 7)
        assert e.value.code == 12
        with pytest.raises(o.OracleError) as e:
            ntt.forward(data, kd, 0, 5, 0, 0, 0, 0)
        assert e.value.code == 13
        with pytest.raises(o.OracleError) as e:
            ntt.forward(data, kd, 0, 6, 0, 4, 2, 0)
        assert e.value.code == 14
        with pytest.raises(o.OracleError) as e:
            ntt.forward(data, kd, 0, 6, 0, 0, 5, 0)
        assert e.value.code == 15


def test_ntt_ext_equals_limbwise(oracle):
    # additive_ntt.rs:137-165 : transforming B128 with B32 twiddles == log_x+2 transform of B32 limbs
    o = oracle
    ntt = o.NTT(5, 10)
    data = o.rand_b128(99, 1 << 6)
    f128 = ntt.forward(data, 7, 1, 5, 0)
    limbs = data.view(np.uint32).reshape(-1)
    f32 = ntt.forward(limbs, 5, 3, 5, 0)
    assert np.array_equal(f128.view(np.uint32).reshape(-1), f32)


def test_fri_fold_matches_definition(oracle):
    # cpu/layer.rs:304-391 with python lerp / inverse butterfly
    o = oracle
    ntt = o.NTT(5, 12)
    rng = random.Random(11)
    for log_len, log_batch, n_ch in [(6, 2, 2), (6, 0, 3), (5, 2, 5), (4, 3, 3), (3, 0, 0)]:
        ch = [rng.getrandbits(128) for _ in range(n_ch)]
        data = o.rand_b128(log_len * 10 + n_ch, 1 << (log_len + log_batch))
        eta = n_ch - log_batch
        n_out = 1 << (log_len - eta)
        got = ints(o, ntt.fri_fold(log_len, log_batch, ch, data, n_out))
        vals = ints(o, data)
        chunk = 1 << n_ch
        for c in range(n_out):
            v = vals[c * chunk:(c + 1) * chunk]
            for r in range(log_batch):
                v = [v[2 * k] ^ o.mul(v[2 * k] ^ v[2 * k + 1], ch[r]) for k in range(len(v) // 2)]
            L, s = log_len, eta
            for r in range(eta):
                nv = []
                for off in range(1 << (s - 1)):
                    t = ntt.get_subspace_eval(L, (c << (s - 1)) | off)
                    u, w = v[2 * off], v[2 * off + 1]
                    w ^= u
                    u ^= o.mul(w, t)
                    nv.append(u ^ o.mul(u ^ w, ch[log_batch + r]))
                v = nv
                L -= 1
                s -= 1
            assert got[c] == v[0]
    with pytest.raises(o.OracleError):
        ntt.fri_fold(4, 2, [1], o.rand_b128(0, 64), 16)


def test_hal_restatements_match_definitions(oracle):
    # fold_left_lerp_inplace with const suffix (fold.rs:648-696) and eq-ind round evals (A.8)
    o = oracle
    rng = random.Random(21)
    n = 5
    full = o.rand_b128(500, 1 << n)
    z = rng.getrandbits(128)
    for prefix, suffix in [(32, 0), (20, rng.getrandbits(128)), (16, 7), (9, rng.getrandbits(128)), (0, 3)]:
        padded = ints(o, full)[:prefix] + [suffix] * (32 - prefix)
        got = ints(o, o.fold_left_lerp_inplace(full[:prefix], prefix, suffix, n, z))
        exp = [padded[i] ^ o.mul(padded[i] ^ padded[16 + i], z) for i in range(16)]
        assert got == exp[: min(prefix, 16)]
        # the dropped tail is again the constant suffix
        assert all(e == suffix for e in exp[min(prefix, 16):])
    # eq-ind round evals vs python loops: sum_i E[i] * C(P_z(i))
    m, nv = 3, 4
    mls = [o.rand_b128(510 + t, 1 << nv) for t in range(m)]
    E = o.rand_b128(520, 1 << (nv - 1))
    comp = [("var", 0), ("var", 1), ("mul", 0, 1), ("var", 2), ("add", 2, 3)]  # x0*x1 + x2
    lead = [("var", 0), ("var", 1), ("mul", 0, 1)]
    pt = rng.getrandbits(128)
    got = o.eq_ind_round_evals(mls, [16] * m, [0] * m, nv, E, [comp], [lead], [1, 2, 3], [0, 0, pt])[0]
    v = [ints(o, x) for x in mls]
    e = ints(o, E)
    r1 = rinf = r3 = 0
    for i in range(8):
        lo = [v[t][i] for t in range(m)]
        hi = [v[t][8 + i] for t in range(m)]
        r1 ^= o.mul(e[i], o.mul(hi[0], hi[1]) ^ hi[2])
        rinf ^= o.mul(e[i], o.mul(hi[0] ^ lo[0], hi[1] ^ lo[1]))
        p = [lo[t] ^ o.mul(lo[t] ^ hi[t], pt) for t in range(m)]
        r3 ^= o.mul(e[i], o.mul(p[0], p[1]) ^ p[2])
    assert got == [r1, rinf, r3]
    h = ints(o, o.fold_partial_eq_ind(E))
    assert h == [e[i] ^ e[4 + i] for i in range(4)]


def test_cpu_baseline_matches_oracle(oracle):
    # the timed CPU arm (AVX-512+GFNI restatement, threaded) is bit-identical to the scalar oracle
    o = oracle
    n = (1 << 10) + 3
    e0, e1 = o.rand_b128(600, n), o.rand_b128(601, n)
    for z in [0, 1, (1 << 128) - 1, 0x2E895399AF449ACE499596F6E5FCCAFA, o.to_ints(o.rand_b128(602, 1))[0]]:
        exp = o.extrapolate_line(e0, e1, z)
        for threads, gfni in [(1, True), (3, True), (2, False)]:
            got, _ = o.cpu_fold(e0, e1, z, threads, gfni)
            assert np.array_equal(got, exp)


def test_cpu_ntt_and_round_eval_arms_match_oracle(oracle):
    """the timed CPU arms of the NTT (port of crates/ntt/src/multithreaded.rs:100-228) and of the bivariate round
    evaluations (fast_compute/src/layer.rs:797-846) are bit-identical to the scalar oracle, GFNI and scalar, any thread count"""
    import random

    o = oracle
    rng = np.random.default_rng(1)
    for (lx, ly, skip, thr) in [(4, 6, 0, 1), (6, 10, 1, 3), (5, 9, 0, 8), (6, 12, 2, 8), (4, 13, 0, 5)]:
        a = rng.integers(0, 1 << 32, size=1 << (lx + ly), dtype=np.uint64).astype(np.uint32)
        exp = o.NTT(5, 16).forward(a, 5, lx, ly, 0, 0, 0, skip)
        for gfni in (True, False):
            assert np.array_equal(o.cpu_ntt_forward(a, lx, ly, skip, 16, thr, gfni), exp), (lx, ly, skip, thr, gfni)
    r = random.Random(2)
    for nv, m, thr in [(4, 3, 1), (13, 4, 3), (14, 5, 8)]:
        mls = [o.rand_b128(50 + t, 1 << nv) for t in range(m)]
        pairs = [(r.randrange(m), r.randrange(m)) for _ in range(5)]
        al = r.getrandbits(128)
        exp = list(o.bivariate_round_evals(mls, nv, pairs, al))
        for gfni in (True, False):
            assert list(o.cpu_bivariate_round_evals(mls, nv, pairs, al, thr, gfni)) == exp, (nv, m, thr, gfni)


def test_cpu_chi_zerocheck_arm_matches_oracle(oracle):
    """the timed CPU arm of the keccak chi zerocheck rounds against the generic eq-ind evaluator restatement"""
    from binius_b200 import ArithCircuit as A

    o = oracle
    n_out, n_b = 5, 7
    for nv, thr in [(3, 1), (11, 3), (12, 8)]:
        cols = [o.rand_b128(900 + t, 1 << nv) for t in range(n_out + n_b)]
        eq = o.rand_b128(77, 1 << (nv - 1))
        v = [A.var(i) for i in range(n_out + n_b)]
        comps = [v[c] - (v[n_out + c % n_b] + (v[n_out + (c + 1) % n_b] - A.one()) * v[n_out + (c + 2) % n_b]) for c in range(n_out)]
        exp = o.eq_ind_round_evals(cols, [len(c) for c in cols], [0] * len(cols), nv, eq, [c.steps for c in comps],
                                   [c.leading_term().steps for c in comps], [1, 2], [0, 0])
        for gfni in (True, False):
            assert o.cpu_chi_round_evals(cols, n_out, n_b, nv, eq, True, thr, gfni) == [list(e) for e in exp], (nv, thr, gfni)


def test_cpu_tensor_expand_arm_matches_oracle(oracle):
    """the timed CPU arm of the eq-indicator expansion against the oracle's tensor_expand"""
    o = oracle
    for log_n, k, thr in [(0, 5, 1), (3, 4, 1), (0, 16, 4), (2, 14, 3)]:
        data = o.rand_b128(40 + log_n, 1 << log_n) if log_n else o.to_arr([1])
        coords = o.to_ints(o.rand_b128(41 + k, k))
        full = np.zeros((1 << (log_n + k), 2), np.uint64)
        full[: 1 << log_n] = data
        exp = o.tensor_expand(full, log_n, coords)
        for gfni in (True, False):
            assert np.array_equal(o.cpu_tensor_expand(data, log_n, coords, thr, gfni), exp), (log_n, k, thr, gfni)
    assert o.cpu_tensor_expand_parallel(12, 0.05, 2)["value"] > 0


def test_cpu_u32add_zerocheck_arm_matches_oracle(oracle):
    """the timed CPU arm of BASELINE config #3 (u32_add zerocheck rounds) against the generic eq-ind evaluator restatement"""
    from binius_b200 import ArithCircuit as A

    o = oracle
    v = [A.var(i) for i in range(5)]
    comps = [(v[0] + v[2]) * (v[1] + v[2]) + v[2] - v[3], v[0] + v[1] + v[2] - v[4]]
    for nv, thr in [(2, 1), (3, 1), (12, 3), (13, 8)]:
        cols = [o.rand_b128(950 + t, 1 << nv) for t in range(5)]
        eq = o.rand_b128(78, 1 << (nv - 1))
        exp = o.eq_ind_round_evals(cols, [len(c) for c in cols], [0] * 5, nv, eq, [c.steps for c in comps],
                                   [c.leading_term().steps for c in comps], [1, 2], [0, 0])
        for gfni in (True, False):
            assert o.cpu_u32add_round_evals(cols, nv, eq, True, thr, gfni) == [exp[0][0], exp[0][1], exp[1][0]], (nv, thr, gfni)
    assert o.cpu_u32add_zerocheck_parallel(10, 0.05, 2)["value"] > 0


def _bitrev(i, n):
    return int(format(i, f"0{n}b")[::-1], 2) if n else 0


def test_low_to_high_restatements(oracle):
    """LowToHigh order (sumcheck_round_calculation.rs:408-504, fold.rs:528-575, common.rs:37-57) against
    definitions: (a) round evals equal the HighToLow ones of the bit-reversed multilinears, (b) the
    regular evaluator satisfies r(0) + r(1) = sum over the hypercube, (c) fold_right_lerp is the
    element-wise lerp of adjacent pairs with the constant suffix, (d) a full LowToHigh sumcheck ends in
    the multilinear evaluation at the challenge point."""
    o = oracle
    rng = random.Random(5)
    n, m = 6, 3
    mls = [o.rand_b128(700 + t, 1 << n) for t in range(m)]
    comp = [("var", 0), ("var", 1), ("mul", 0, 1), ("var", 2), ("add", 2, 3)]  # x*y + w
    lead = [("var", 0), ("var", 1), ("mul", 0, 1)]
    eq = o.rand_b128(710, 1 << (n - 1))
    z3 = rng.getrandbits(128)
    codes, pts = [1, 2, 3], [0, 0, z3]
    lens, sfx = [1 << n] * m, [0] * m
    lo2hi = o.sumcheck_round_evals(0, mls, lens, sfx, n, eq, [comp], [lead], codes, pts)
    # (a) bit-reversed inputs in the other order; eq-ind indices follow the remaining n-1 variables
    rev = [np.stack([x[_bitrev(i, n)] for i in range(1 << n)]) for x in mls]
    eq_rev = np.stack([eq[_bitrev(i, n - 1)] for i in range(1 << (n - 1))])
    assert o.sumcheck_round_evals(1, rev, lens, sfx, n, eq_rev, [comp], [lead], codes, pts) == lo2hi
    assert o.sumcheck_round_evals(1, mls, lens, sfx, n, eq, [comp], [lead], codes, pts) == o.eq_ind_round_evals(mls, lens, sfx, n, eq, [comp], [lead], codes, pts)
    # (b) regular evaluator: r(0) (finite point 0) + r(1) = sum_x C(M(x))
    for order in (0, 1):
        r = o.sumcheck_round_evals(order, mls, lens, sfx, n, None, [comp], [lead], [1, 3], [0, 0])[0]
        total = o.sum_composition_evals(mls, comp, 1, 0)
        assert r[0] ^ r[1] == total
    # (c) fold_right_lerp with an odd stored prefix and a constant suffix
    z = rng.getrandbits(128)
    for plen, s in ((64, 0), (37, rng.getrandbits(128)), (1, 7), (0, 9)):
        x = mls[0][:plen]
        got = ints(o, o.fold_right_lerp(x, s, z))
        xi = ints(o, x) + [s]
        assert got == [xi[2 * i] ^ o.mul(xi[2 * i] ^ xi[2 * i + 1], z) for i in range((plen + 1) // 2)]
    assert ints(o, o.fold_partial_eq_ind_low_to_high(eq)) == [a ^ b for a, b in zip(ints(o, eq)[0::2], ints(o, eq)[1::2])]
    # (d) fold all variables low-to-high: the result is the multilinear evaluated at the challenges
    ch = [rng.getrandbits(128) for _ in range(n)]
    cur = mls[1]
    for c in ch:
        cur = o.fold_right_lerp(cur, 0, c)
    basis = o.tensor_expand(o.to_arr([1] + [0] * ((1 << n) - 1)), 0, ch)
    assert ints(o, cur) == [o.inner_product(mls[1], 7, basis)]
