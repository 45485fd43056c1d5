import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import binius_b200
hal = binius_b200.B200Layer(0)
k = int(os.environ.get("K", "22"))
dev = hal.dev_alloc(1 << k)
rr = random.Random(7)
coords = [rr.getrandbits(128) for _ in range(k)]
hal.fill(dev.slice(0, 1), 1)
for _ in range(3):
    hal.execute(lambda ex: (ex.tensor_expand(0, coords, dev), [])[1])
hal.sync()
