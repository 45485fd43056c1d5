"""GPU parity tests of the zerocheck univariate-skip round (b200_zerocheck_univariate_evals) against the
oracle (oracle/univariate.c), through the C ABI via the Python mirror of
`zerocheck_univariate_evals` (core/src/protocols/sumcheck/prove/univariate.rs:235-500)."""
import random

import numpy as np
import pytest

from test_oracle_univariate import pack_scalars, zero_product_columns

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hal():
    import binius_b200

    layer = binius_b200.B200Layer(0)
    yield layer
    layer.close()


def _product(ix):
    from binius_b200 import ArithCircuit as A

    acc = A.var(ix[0])
    for i in ix[1:]:
        acc = acc * A.var(i)
    return acc


def _run(hal, oracle, cols, levels, n_vars, skip, comps, max_domain, challenges):
    from binius_b200.hal import B200Backend, TransparentMultilinear, _degree, zerocheck_univariate_evals

    be = B200Backend(hal)
    packed = [oracle.to_arr(pack_scalars(c, l)) for c, l in zip(cols, levels)]
    mls = [TransparentMultilinear(hal.to_device(p), l, n_vars) for p, l in zip(packed, levels)]
    out = zerocheck_univariate_evals(be, mls, comps, challenges, skip, max_domain)
    eq = oracle.tensor_expand(oracle.to_arr([1] + [0] * ((1 << len(challenges)) - 1)), 0, challenges)
    assert np.array_equal(hal.to_host(out.partial_eq_ind_evals), eq)
    exp = oracle.zerocheck_univariate_evals_reference(packed, levels, n_vars, skip, eq, [list(c.steps) for c in comps],
                                                      [_degree(c) for c in comps], max_domain)
    assert out.skip_rounds == skip and out.remaining_rounds == n_vars - skip and out.max_domain_size == max_domain
    return out.round_evals, exp, packed, eq


@pytest.mark.parametrize("skip", [0, 1, 2, 3, 4, 5])
def test_reference_test_setting(hal, oracle, skip):
    """univariate.rs:806-915: n_vars = 7, zero-product B1 multilinears of degree 2, 3 and 4,
    max_domain_size = domain_size(5, skip); compared with the definition (the reference's naive check)."""
    rng = random.Random(skip)
    n_vars = 7
    cols = zero_product_columns(rng, n_vars, 2) + zero_product_columns(rng, n_vars, 3) + zero_product_columns(rng, n_vars, 4)
    comps = [_product([0, 1]), _product([2, 3, 4]), _product([5, 6, 7, 8])]
    ch = [rng.getrandbits(128) for _ in range(n_vars - skip)]
    got, exp, packed, eq = _run(hal, oracle, cols, [0] * 9, n_vars, skip, comps, 5 << skip, ch)
    assert all(len(v) == 4 << skip for v in got)
    assert got == exp
    naive = oracle.zerocheck_univariate_evals(packed, [0] * 9, n_vars, skip, eq, [list(c.steps) for c in comps], 5 << skip)
    assert got == naive


@pytest.mark.parametrize("levels,skip,n_vars", [((0, 3, 0, 3), 3, 8), ((0, 4, 3, 0), 2, 7), ((5, 0, 3, 4), 4, 8), ((7, 0, 6, 3), 1, 6),
                                                ((0, 0, 0, 0), 6, 10), ((3, 3, 0, 0), 7, 9)])
def test_mixed_levels_random_instances(hal, oracle, levels, skip, n_vars):
    """Random (unsatisfied) instances over mixed sub-fields, constants and powers in the compositions, a
    composition of lower degree than the maximum (extended the reference's way) and a linear one (all zero)."""
    from binius_b200 import ArithCircuit as A

    rng = random.Random(hash((levels, skip)) & 0xFFFF)
    cols = [[rng.getrandbits(1 << l) for _ in range(1 << n_vars)] for l in levels]
    base = max(max(levels), 3)
    cst = rng.getrandbits(1 << base) | 1
    x, y, z, w = (A.var(i) for i in range(4))
    comps = [(x + z) * (y + w) + A.constant(cst) * z, x + y + w, (x * y + z).pow(2) if skip < 6 else x * y, x * w * (y + A.constant(1))]
    from binius_b200.hal import _degree

    max_deg = max(_degree(c) for c in comps)
    max_domain = min(256, (max_deg << skip) + (3 if skip < 5 else 0))
    if max_deg << skip > 256:
        comps = comps[:2]
        max_domain = 256 if skip == 7 else 2 << skip
    ch = [rng.getrandbits(128) for _ in range(n_vars - skip)]
    got, exp, _, _ = _run(hal, oracle, cols, list(levels), n_vars, skip, comps, max_domain, ch)
    assert got == exp
    assert all(v == 0 for v in got[1])


def test_many_subcubes_keccak_like_shape(hal, oracle):
    """B1 columns, skip = 6 (64-point sub-cubes), degree-2 chi-like constraints, 2^14 rows: many sub-cubes per
    CTA lane and several CTAs per point block."""
    from binius_b200 import ArithCircuit as A

    rng = random.Random(77)
    n_vars, skip, m = 14, 6, 6
    cols = [[rng.getrandbits(1) for _ in range(1 << n_vars)] for _ in range(m)]
    v = [A.var(i) for i in range(m)]
    comps = [(v[0] + A.constant(1)) * v[1] + v[2] + v[3], v[4] * v[5] + v[0], v[1] * v[3] + v[2] * v[5]]
    ch = [rng.getrandbits(128) for _ in range(n_vars - skip)]
    got, exp, _, _ = _run(hal, oracle, cols, [0] * m, n_vars, skip, comps, 128, ch)
    assert all(len(r) == 64 for r in got)
    assert got == exp


def test_linear_composition_at_full_size(hal, oracle):
    """Size-independent property at 2^20 rows (the keccak 2^18 column length after packing is 2^27 bits; this
    is the largest size whose host-side weights are built in seconds): for a composition equal to var(0),
    R[i] = <M_0, L(x_i) (x) eq>, checked with the device inner product on the tensor built on the host.
    The composition is written x0 + x1*x1 + x1*x1 (= x0 in characteristic 2) so that its declared degree is 2
    and it is evaluated at 2^skip real points instead of being short-circuited to zero."""
    from binius_b200 import ArithCircuit as A, SubfieldSlice
    from binius_b200.hal import B200Backend, TransparentMultilinear, zerocheck_univariate_evals

    n_vars, skip = 20, 6
    K = 1 << skip
    be = B200Backend(hal)
    mls = [TransparentMultilinear(hal.to_device(oracle.rand_b128(41 + j, 1 << (n_vars - 7))), 0, n_vars) for j in range(2)]
    rng = random.Random(3)
    ch = [rng.getrandbits(128) for _ in range(n_vars - skip)]
    x0, x1 = A.var(0), A.var(1)
    out = zerocheck_univariate_evals(be, mls, [x0 + x1 * x1 + x1 * x1], ch, skip, 2 * K)
    eq = hal.to_host(out.partial_eq_ind_evals)
    for i in (0, 37, K - 1):
        lag = oracle.lagrange_evals(skip, K + i)
        w = np.zeros((len(eq), K, 2), np.uint64)  # w[s][t] = eq[s] * L_t(x_i)
        for t in range(K):
            w[:, t, :] = oracle.mul_vec(eq, oracle.to_arr([lag[t]] * len(eq)))
        wd = hal.to_device(w.reshape(-1, 2))
        got_ip = hal.execute(lambda ex: [ex.inner_product(SubfieldSlice(mls[0].evals, 0), wd)])[0]
        hal.dev_free(wd)
        assert out.round_evals[0][i] == got_ip


def test_error_behaviour(hal, oracle):
    from binius_b200 import ArithCircuit as A, InputValidation
    from binius_b200.hal import B200Backend, TransparentMultilinear, zerocheck_univariate_evals

    be = B200Backend(hal)
    n_vars = 6
    ml = TransparentMultilinear(hal.to_device(oracle.rand_b128(1, 1)), 0, n_vars)
    comp = A.var(0) * A.var(0)
    with pytest.raises(InputValidation):  # TooManySkippedRounds
        zerocheck_univariate_evals(be, [ml], [comp], [], 7, 256)
    with pytest.raises(InputValidation):  # IncorrectZerocheckChallengesLength
        zerocheck_univariate_evals(be, [ml], [comp], [1, 2, 3], 2, 8)
    with pytest.raises(InputValidation):  # LagrangeDomainTooSmall
        zerocheck_univariate_evals(be, [ml], [comp], [1, 2, 3, 4], 2, 7)
    with pytest.raises(InputValidation):  # DomainSizeTooLarge
        zerocheck_univariate_evals(be, [ml], [comp], [1, 2, 3, 4], 2, 300)
    # the C ABI validates on its own as well
    import ctypes as C

    L = hal
    eq = be.tensor_product_full_query([1, 2, 3, 4])
    ptrs = (C.c_void_p * 1)(ml.evals.ptr)
    lv = (C.c_uint32 * 1)(0)
    ce = (C.c_void_p * 1)(be._compiled(comp)[0].handle.value)
    dg = (C.c_uint32 * 1)(2)
    out = (C.c_uint64 * 64)()
    assert L._lib.b200_zerocheck_univariate_evals(L._ctx, ptrs, lv, 1, n_vars, 2, eq.ptr, 15, ce, dg, 1, 8, out) == 1
    assert L._lib.b200_zerocheck_univariate_evals(L._ctx, ptrs, lv, 1, n_vars, 2, eq.ptr, 16, ce, dg, 1, 7, out) == 1
    assert L._lib.b200_zerocheck_univariate_evals(L._ctx, ptrs, lv, 1, n_vars, 7, eq.ptr, 16, ce, dg, 1, 8, out) == 1
    lv[0] = 2
    assert L._lib.b200_zerocheck_univariate_evals(L._ctx, ptrs, lv, 1, n_vars, 2, eq.ptr, 16, ce, dg, 1, 8, out) == 1
    # max_domain_size == 2^skip: nothing to evaluate
    lv[0] = 0
    dg[0] = 1
    lin = (C.c_void_p * 1)(be._compiled(A.var(0))[0].handle.value)
    assert L._lib.b200_zerocheck_univariate_evals(L._ctx, ptrs, lv, 1, n_vars, 2, eq.ptr, 16, lin, dg, 1, 4, out) == 0


@pytest.mark.parametrize("generic", [False, True])
def test_many_columns_at_the_reference_skip(hal, oracle, generic):
    """The reference's choice for degree-2 constraints over B8 is skip_rounds = 7 (constraint_system/verify.rs:271-294).
    80 columns at skip 7 exceed one CTA's shared memory: the default path runs the fast kernel per composition range
    (uni_split.hpp); `generic` forces the generic kernel on the same shape."""
    from binius_b200 import ArithCircuit as A

    hal.set_tuning("uni_generic", int(generic))
    rng = random.Random(123)
    n_vars, skip, m = 9, 7, 80
    cols = [[rng.getrandbits(1) for _ in range(1 << n_vars)] for _ in range(m)]
    v = [A.var(i) for i in range(m)]
    comps = [v[(2 * c) % m] * v[(2 * c + 1) % m] + v[(2 * c + 5) % m] + v[(2 * c + 11) % m] for c in range(36)]
    comps.append(v[3] * v[70] + A.constant(0x53) * v[40] + A.constant(1))
    ch = [rng.getrandbits(128) for _ in range(n_vars - skip)]
    try:
        got, exp, _, _ = _run(hal, oracle, cols, [0] * m, n_vars, skip, comps, 256, ch)
    finally:
        hal.set_tuning("uni_generic", 0)
    assert got == exp


@pytest.mark.parametrize("skip,log_chunks", [(7, 2), (5, 3), (3, 1)])
def test_streamed_round_equals_the_resident_one(hal, oracle, skip, log_chunks):
    """zerocheck_univariate_evals_streamed: the witness starts in pinned host memory and is uploaded chunk by chunk on the
    side stream while the previous chunk is evaluated; values, eq-indicator and the uploaded columns must equal the
    resident call's (mixed B1 / B8 columns, degree 2 and 3 compositions -> the domain extension runs per chunk)."""
    from binius_b200 import ArithCircuit as A
    from binius_b200.hal import B200Backend, TransparentMultilinear, zerocheck_univariate_evals, zerocheck_univariate_evals_streamed

    be = B200Backend(hal)
    rng = random.Random(skip)
    n_vars = 13
    levels = [0, 0, 0, 3, 0]
    host_cols = []
    for j, lvl in enumerate(levels):
        h = hal.host_alloc((1 << n_vars << lvl) // 128)
        h[:] = oracle.rand_b128(7000 + j, len(h))
        host_cols.append(h)
    v = [A.var(i) for i in range(5)]
    comps = [v[0] * v[1] + v[2] + v[4], v[3] * v[4] + A.constant(0x35) * v[0]]
    if skip < 7:  # degree 3 at skip 7 would need a 384-point domain (> 256 = |B8|)
        comps.append(v[0] * v[2] * v[4] + v[1])
    max_domain = min(3 << skip, 256)
    ch = [rng.getrandbits(128) for _ in range(n_vars - skip)]
    resident = [TransparentMultilinear(hal.to_device(h.copy()), lvl, n_vars) for h, lvl in zip(host_cols, levels)]
    exp = zerocheck_univariate_evals(be, resident, comps, ch, skip, max_domain)
    dst = [TransparentMultilinear(hal.dev_alloc(len(h)), lvl, n_vars) for h, lvl in zip(host_cols, levels)]
    for d in dst:
        hal.fill(d.evals, 0)
    hal.sync()
    got = zerocheck_univariate_evals_streamed(be, host_cols, dst, comps, ch, skip, max_domain, log_chunks)
    assert got.round_evals == exp.round_evals
    assert np.array_equal(hal.to_host(got.partial_eq_ind_evals), hal.to_host(exp.partial_eq_ind_evals))
    for d, h in zip(dst, host_cols):
        assert np.array_equal(hal.to_host(d.evals), h)


def test_streamed_round_from_a_witness_arena_with_the_linear_route(hal, oracle):
    """Columns at one stride in host and device memory: every chunk is ONE pitched copy; at skip 7 with >= 2^12 sub-cubes
    per chunk the linear monomials of every chunk take the tensor-core route.  Against the resident call with the
    route switched off."""
    from binius_b200 import ArithCircuit as A
    from binius_b200.hal import B200Backend, TransparentMultilinear, zerocheck_univariate_evals, zerocheck_univariate_evals_streamed

    be = B200Backend(hal)
    n_vars, skip, m, log_chunks = 21, 7, 6, 1
    words = (1 << n_vars) // 128
    h_arena = hal.host_alloc(m * words)
    h_arena[:] = oracle.rand_b128(7100, m * words)
    host_cols = [h_arena[j * words:(j + 1) * words] for j in range(m)]
    v = [A.var(i) for i in range(m)]
    comps = [v[0] * v[1] + v[2] + v[3], v[4] * v[0] + v[5] + A.constant(0x1D)]
    ch = [random.Random(71).getrandbits(128) for _ in range(n_vars - skip)]
    resident = [TransparentMultilinear(hal.to_device(np.array(h)), 0, n_vars) for h in host_cols]
    hal.set_tuning("uni_linear", 0)
    try:
        exp = zerocheck_univariate_evals(be, resident, comps, ch, skip, 256)
    finally:
        hal.set_tuning("uni_linear", 1)
    d_arena = hal.dev_alloc(m * words)
    hal.fill(d_arena, 0)
    dst = [TransparentMultilinear(d_arena.slice(j * words, (j + 1) * words), 0, n_vars) for j in range(m)]
    hal.sync()
    got = zerocheck_univariate_evals_streamed(be, host_cols, dst, comps, ch, skip, 256, log_chunks)
    assert got.round_evals == exp.round_evals
    assert np.array_equal(hal.to_host(d_arena), np.array(h_arena))


# ---- the round in two halves: prepare (no challenge) + finish (b200_zerocheck_univariate_prepare / _finish) ----
def _two_halves(hal, oracle, cols, levels, n_vars, skip, comps, max_domain, challenges, streamed, log_chunks=2):
    """prepare + finish against the one-call round on the same device columns; returns (two halves, one call, prepared)"""
    from binius_b200.hal import (B200Backend, TransparentMultilinear, zerocheck_univariate_evals, zerocheck_univariate_finish,
                                 zerocheck_univariate_prepare)

    be = B200Backend(hal)
    packed = [oracle.to_arr(pack_scalars(c, l)) if not isinstance(c, np.ndarray) else c for c, l in zip(cols, levels)]
    if streamed:
        mls = [TransparentMultilinear(hal.dev_alloc(len(p)), l, n_vars) for p, l in zip(packed, levels)]
        for ml in mls:
            hal.fill(ml.evals, 0x5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A)  # the upload must overwrite this
        hosts = []
        for p in packed:
            h = hal.host_alloc(len(p))
            h[:] = p
            hosts.append(h)
        hal.sync()
        prep = zerocheck_univariate_prepare(be, mls, comps, skip, max_domain, host_columns=hosts, log_chunks=log_chunks)
        for ml, p in zip(mls, packed):
            assert np.array_equal(hal.to_host(ml.evals), p), "columns resident after the streamed prepare"
    else:
        mls = [TransparentMultilinear(hal.to_device(p), l, n_vars) for p, l in zip(packed, levels)]
        arena = hal.dev_alloc(1 << 16)  # the store as a slice of caller-owned memory (4 x what the largest case here needs)
        prep = zerocheck_univariate_prepare(be, mls, comps, skip, max_domain, arena_store=arena)
        assert not prep.owns_store
    prepared = prep.prepared
    out = zerocheck_univariate_finish(be, prep, challenges)
    prep.release(be)
    ref = zerocheck_univariate_evals(be, mls, comps, challenges, skip, max_domain)
    assert np.array_equal(hal.to_host(out.partial_eq_ind_evals), hal.to_host(ref.partial_eq_ind_evals))
    return out.round_evals, ref.round_evals, prepared, packed


@pytest.mark.parametrize("skip,n_vars,streamed", [(2, 9, False), (4, 11, True), (6, 12, False), (7, 13, True), (5, 5, False), (4, 8, True)])
def test_prepare_finish_equals_the_one_call_round(hal, oracle, skip, n_vars, streamed):
    """B1/B8 columns, degree-2 monomials with constants, a composition of lower degree (extended the reference's way) and a
    linear one; a ragged last batch (n_vars - skip < 3) at skip 5; at (4, 8) the requested chunking (4 chunks of 4 sub-cubes)
    is coarsened to whole batches of the store; vs the one-call round AND vs the oracle."""
    from binius_b200 import ArithCircuit as A
    from binius_b200.hal import _degree

    rng = random.Random(100 * skip + n_vars)
    levels = [0, 3, 0, 0, 3]
    cols = [[rng.getrandbits(1 << l) for _ in range(1 << n_vars)] for l in levels]
    x, y, z, w, u = (A.var(i) for i in range(5))
    comps = [(x + z) * (y + w) + A.constant(rng.getrandbits(8) | 1) * z, x * w + u, x + y + w, u * y + A.constant(0x1D) * (x * z) + A.constant(7)]
    ch = [rng.getrandbits(128) for _ in range(n_vars - skip)]
    got, ref, prepared, packed = _two_halves(hal, oracle, cols, levels, n_vars, skip, comps, 2 << skip, ch, streamed)
    assert prepared
    assert got == ref
    eq = oracle.tensor_expand(oracle.to_arr([1] + [0] * ((1 << len(ch)) - 1)), 0, ch)
    exp = oracle.zerocheck_univariate_evals_reference(packed, levels, n_vars, skip, eq, [list(c.steps) for c in comps], [_degree(c) for c in comps], 2 << skip)
    assert got == exp


def test_prepare_declines_shapes_outside_the_fast_path(hal, oracle):
    """a B32 column / a cubic composition: prepare stores nothing and finish runs the whole round itself"""
    from binius_b200 import ArithCircuit as A

    rng = random.Random(9)
    n_vars, skip = 8, 3
    for levels, comps in (((0, 5, 0), [A.var(0) * A.var(1) + A.var(2)]), ((0, 0, 0), [A.var(0) * A.var(1) * A.var(2)]),
                          ((0, 3, 0), [A.var(0) + A.var(1) + A.var(2)])):  # (linear only: nothing to extrapolate)
        cols = [[rng.getrandbits(1 << l) for _ in range(1 << n_vars)] for l in levels]
        ch = [rng.getrandbits(128) for _ in range(n_vars - skip)]
        got, ref, prepared, _ = _two_halves(hal, oracle, cols, levels, n_vars, skip, comps, 3 << skip, ch, False)
        assert not prepared
        assert got == ref


def test_prepare_finish_keccak_shape_2pow20_rows(hal, oracle):
    """153 B1 columns, the 75 chi constraints, skip 7, 2^13 sub-cubes (the linear-monomial route is on): streamed prepare in 4
    chunks + finish vs the one-call round and vs the CPU arm"""
    from test_gpu_baseline_sizes import keccak_chi_compositions

    n_vars, skip, m = 20, 7, 153
    rng = random.Random(77)
    cols = [oracle.rand_b128(4530 + j, (1 << n_vars) // 128) for j in range(m)]
    comps = keccak_chi_compositions()
    ch = [rng.getrandbits(128) for _ in range(n_vars - skip)]
    got, ref, prepared, _ = _two_halves(hal, oracle, cols, [0] * m, n_vars, skip, comps, 256, ch, True, log_chunks=2)
    assert prepared and got == ref
    eq = oracle.tensor_expand(oracle.to_arr([1] + [0] * ((1 << len(ch)) - 1)), 0, ch)
    exp, _ = oracle.cpu_univariate_b1(cols, n_vars, skip, eq, [list(c.steps) for c in comps], 1 << skip)
    assert got == exp


def test_prepare_finish_error_classes(hal, oracle):
    """store too small -> InputValidation from prepare and from finish; finish on a shape prepare does not cover ->
    InputValidation; wrong number of challenges -> IncorrectZerocheckChallengesLength (Python mirror)"""
    import ctypes as C

    import binius_b200
    from binius_b200 import ArithCircuit as A
    from binius_b200.hal import (B200Backend, TransparentMultilinear, _uni_call_args, zerocheck_univariate_finish, zerocheck_univariate_prepare)

    be = B200Backend(hal)
    rng = random.Random(5)
    n_vars, skip = 10, 4
    cols = [[rng.getrandbits(1) for _ in range(1 << n_vars)] for _ in range(3)]
    mls = [TransparentMultilinear(hal.to_device(oracle.to_arr(pack_scalars(c, 0))), 0, n_vars) for c in cols]
    comps = [A.var(0) * A.var(1) + A.var(2)]
    L = hal
    ptrs, lvls, cps, degs = _uni_call_args(be, mls, comps)
    need = int(L._lib.b200_zerocheck_univariate_store_elems(n_vars, skip, degs, 1))
    assert need == ((1 << (n_vars - skip)) * 1 * (1 << skip)) // 16
    small = hal.dev_alloc(max(need - 1, 1))
    done = C.c_uint32()
    with pytest.raises(binius_b200.InputValidation):
        L._check(L._lib.b200_zerocheck_univariate_prepare(L._ctx, None, ptrs, lvls, 3, n_vars, skip, cps, degs, 1, 2 << skip, 0, small.ptr, C.c_uint64(small.len()), C.byref(done)))
    eq = be.tensor_product_full_query([rng.getrandbits(128) for _ in range(n_vars - skip)])
    out = (C.c_uint64 * (2 * (1 << skip)))()
    with pytest.raises(binius_b200.InputValidation):
        L._check(L._lib.b200_zerocheck_univariate_finish(L._ctx, ptrs, lvls, 3, n_vars, skip, eq.ptr, C.c_uint64(eq.len()), cps, degs, 1, 2 << skip, small.ptr,
                                                         C.c_uint64(small.len()), out))
    # a cubic composition is not a prepared shape: the C entry point refuses, the mirror falls back to the one-call round
    cubic = [A.var(0) * A.var(1) * A.var(2)]
    p2, l2, c2, d2 = _uni_call_args(be, mls, cubic)
    big = hal.dev_alloc(4 * need + 16)
    out3 = (C.c_uint64 * (2 * (2 << skip)))()
    with pytest.raises(binius_b200.InputValidation):
        L._check(L._lib.b200_zerocheck_univariate_finish(L._ctx, p2, l2, 3, n_vars, skip, eq.ptr, C.c_uint64(eq.len()), c2, d2, 1, 3 << skip, big.ptr, C.c_uint64(big.len()), out3))
    prep = zerocheck_univariate_prepare(be, mls, comps, skip, 2 << skip)
    assert prep.prepared
    with pytest.raises(binius_b200.InputValidation):
        zerocheck_univariate_finish(be, prep, [1, 2, 3])
    prep.release(be)
    for d in (small, big, eq):
        hal.dev_free(d)
