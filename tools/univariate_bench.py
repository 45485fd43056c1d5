"""Timing of the zerocheck univariate-skip round (b200_zerocheck_univariate_evals) at the keccak shape
(153 B1 columns, 75 degree-2 chi constraints, skip_rounds = 7 (constraint_system/verify.rs:271-294),
max_domain_size = 256) on synthetic columns.  `python tools/univariate_bench.py [n_vars] [m] [n_comp]`;
n_vars = 27 is the 2^18-permutation trace (16 MiB per column).  Run on a B200."""
import ctypes as C
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import binius_b200
from binius_b200 import ArithCircuit as A
from binius_b200.hal import B200Backend, TransparentMultilinear, zerocheck_univariate_evals

n_vars = int(sys.argv[1]) if len(sys.argv) > 1 else 24
m = int(sys.argv[2]) if len(sys.argv) > 2 else 153
n_comp = int(sys.argv[3]) if len(sys.argv) > 3 else 75
skip = 7
hal = binius_b200.B200Layer(0)
be = B200Backend(hal)
words = 1 << (n_vars - 7)
arena = hal.dev_alloc(m * words)
hal.fill(arena, 0x0123456789ABCDEF0F1E2D3C4B5A6978)
rng = random.Random(1)  # (the kernel's work does not depend on the column values)
mls = [TransparentMultilinear(arena.slice(j * words, (j + 1) * words), 0, n_vars) for j in range(m)]
comps = []
for c in range(n_comp):  # keccak chi constraints out - (b0 + (b1 - 1) * b2) (stacked.rs:318-366) when the shape allows
    if m >= 150 and n_comp <= 75:
        batch, xy = c // 25, c % 25
        x, y = xy % 5, xy // 5
        b0, b1, b2 = (A.var(75 + 25 * batch + (x + k) % 5 + 5 * y) for k in range(3))
        comps.append(A.var(c) - (b0 + (b1 - A.one()) * b2))
    else:
        a, b, d, e = (A.var((2 * c + o) % m) for o in (0, 1, 5, 11))
        comps.append(a * b + d + e)
ch = [rng.getrandbits(128) for _ in range(n_vars - skip)]
if len(sys.argv) > 4 and sys.argv[4] == "split":  # the round in two halves (b200_zerocheck_univariate_prepare / _finish)
    from binius_b200.hal import zerocheck_univariate_finish, zerocheck_univariate_prepare

    for rep in range(3):
        hal.sync()
        t0 = time.perf_counter()
        prep = zerocheck_univariate_prepare(be, mls, comps, skip, 2 << skip)
        hal.sync()
        t1 = time.perf_counter()
        out = zerocheck_univariate_finish(be, prep, ch)
        t2 = time.perf_counter()
        print(f"n_vars={n_vars} m={m} compositions={n_comp}: prepare {1e3 * (t1 - t0):.2f} ms (resident columns), finish {1e3 * (t2 - t1):.2f} ms, "
              f"store {prep.store.len() * 16 / 1e9:.2f} GB, nonzero={sum(1 for r in out.round_evals for v in r if v)}")
        hal.dev_free(out.partial_eq_ind_evals)
        prep.release(be)
    sys.exit(0)
for rep in range(3):
    hal.sync()
    t0 = time.perf_counter()
    out = zerocheck_univariate_evals(be, mls, comps, ch, skip, 2 << skip)  # synchronous (returns host values)
    dt = time.perf_counter() - t0
    n_sub = 1 << (n_vars - skip)
    print(f"n_vars={n_vars} m={m} compositions={n_comp}: {dt * 1e3:.2f} ms per call "
          f"({n_sub * m / dt / 1e9:.2f} G sub-cube columns/s, {m * words * 16 / dt / 1e9:.1f} GB/s of column data), "
          f"nonzero={sum(1 for r in out.round_evals for v in r if v)}")
    hal.dev_free(out.partial_eq_ind_evals)
