"""profiling stub: the projection after the univariate round -- 153 fold_right calls of a B1 column (2^20 words) by one
128-coefficient query, batched by the library into one k_fold_right_lut_multi launch"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import binius_b200

hal = binius_b200.B200Layer(0)
n, m = 1 << 20, 153
q = hal.dev_alloc(128)
hal.fill(q, 0x0123456789ABCDEF0F1E2D3C4B5A6978)
cols = hal.dev_alloc(m * n)
hal.fill(cols, 0x8796A5B4C3D2E1F10F1E2D3C4B5A6978)
out = hal.dev_alloc(m * n)
S = binius_b200.SubfieldSlice
for rep in range(3):
    hal.execute(lambda ex: [ex.fold_right(S(cols.slice(j * n, (j + 1) * n), 0), q, out.slice(j * n, (j + 1) * n)) for j in range(m)] and [])
hal.sync()
