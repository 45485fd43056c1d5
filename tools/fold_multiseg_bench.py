"""Fold-high throughput for several multilinears per call (b200_fold_multilinears_high_to_low: one launch per
48 segments) and for small sizes: shows the fixed cost per launch (~8 us: launch + table build + pipeline fill)
and that from ~2^25 coefficients per call the kernel runs at the measured HBM peak.  Run on a B200."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import binius_b200
hal = binius_b200.B200Layer(0)
nv = 20
M = 144
arena = hal.dev_alloc(M << nv)
hal.fill(arena, 0x1234567)
a, b = C.c_void_p(), C.c_void_p()
hal._lib.b200_event_create(hal._ctx, C.byref(a)); hal._lib.b200_event_create(hal._ctx, C.byref(b))
z = (C.c_uint64 * 2)(5, 9)
def run(m, n_vars):
    ptrs = (C.c_void_p * m)(*[arena.ptr + 16 * (i << n_vars) for i in range(m)])
    lens = (C.c_uint64 * m)(*[1 << n_vars] * m)
    sfx = (C.c_uint64 * (2 * m))()
    nl = (C.c_uint64 * m)()
    for _ in range(3):
        hal._check(hal._lib.b200_fold_multilinears_high_to_low(hal._ctx, ptrs, m, n_vars, lens, sfx, z, nl))
    hal.sync()
    hal._lib.b200_event_record(hal._ctx, a)
    reps = 10
    for _ in range(reps):
        hal._check(hal._lib.b200_fold_multilinears_high_to_low(hal._ctx, ptrs, m, n_vars, lens, sfx, z, nl))
    hal._lib.b200_event_record(hal._ctx, b)
    ms = C.c_float(); hal._lib.b200_event_elapsed_ms(hal._ctx, a, b, C.byref(ms))
    t = ms.value / reps
    byts = 24.0 * m * (1 << n_vars)
    print(f"m={m:4d} n_vars={n_vars}: {t:.4f} ms  {byts / t / 1e6:.0f} GB/s")
run(1, 24); run(1, 20); run(16, 20); run(48, 20); run(96, 20); run(144, 20); run(48, 16); run(144, 14)
