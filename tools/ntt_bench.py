"""NTT timing on device-resident data (CUDA events on the library's stream): S1/S2/S3 of BASELINE config #2 and
the keccak RS-encode shape, for each kernel selection (`ntt` tuning key: 0 look-up tables, 1 bit-sliced only)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import binius_b200
from binius_b200 import NTTShape

hal = binius_b200.B200Layer(0)
lib, ctx = hal._lib, hal._ctx
modes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["0", "1"])]
big = "--big" in sys.argv


def timed(fn, reps=10):
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.b200_event_create(ctx, C.byref(e0))
    lib.b200_event_create(ctx, C.byref(e1))
    for _ in range(3):
        fn()
    lib.b200_event_record(ctx, e0)
    for _ in range(reps):
        fn()
    lib.b200_event_record(ctx, e1)
    ms = C.c_float()
    lib.b200_event_elapsed_ms(ctx, e0, e1, C.byref(ms))
    return ms.value / reps


out = {}
shapes = {"S1": (24, 6, 18, 0, 1), "S2": (24, 0, 24, 0, 0), "S3": (24, 0, 16, 8, 0)}
if big:
    shapes["rs_encode_2^30"] = (30, 6, 24, 0, 1)
ntt = binius_b200.B200AdditiveNTT(hal, 5, 26)
dev = hal.dev_alloc(1 << (28 if big else 22))
for mode in modes:
    hal.set_tuning("ntt", mode)
    for cc, cw in ([(c, w) for c in (6, 7) for w in (8, 16)] if mode == 0 and "--cc" in sys.argv else [(7, 16)]):
        hal.set_tuning("ntt_log_cc", cc)
        hal.set_tuning("ntt_cw", cw)
        for name, (ln, lx, ly, lz, skip) in shapes.items():
            l0 = hal.launch_count()
            f = timed(lambda: ntt.forward_device(dev.ptr, 5, 1 << ln, NTTShape(lx, ly, lz), 0, 0, skip))
            launches = (hal.launch_count() - l0) // 13
            i = timed(lambda: ntt.inverse_device(dev.ptr, 5, 1 << ln, NTTShape(lx, ly, lz), 0, 0, skip))
            out[f"mode{mode}_cc{cc}_cw{cw}_{name}"] = {"fwd_ms": round(f, 4), "inv_ms": round(i, 4), "passes": launches}
print(json.dumps(out, indent=1))
