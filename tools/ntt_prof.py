import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import binius_b200
from binius_b200 import NTTShape
hal = binius_b200.B200Layer(0)
ntt = binius_b200.B200AdditiveNTT(hal, 5, 24)
dev = hal.dev_alloc(1 << 22)
shape = sys.argv[1] if len(sys.argv) > 1 else "S1"
lx, ly, lz, skip = {"S1": (6, 18, 0, 1), "S2": (0, 24, 0, 0), "S3": (0, 16, 8, 0)}[shape]
for _ in range(3):
    ntt.forward_device(dev.ptr, 5, 1 << 24, NTTShape(lx, ly, lz), 0, 0, skip)
hal.sync()
