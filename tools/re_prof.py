import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import binius_b200
hal = binius_b200.B200Layer(0)
nv, m = int(os.environ.get("NV", "20")), 8
dev = hal.dev_alloc(m << nv)
hal.fill(dev, 0x123456789ABCDEF0FEDCBA9876543211)
sub = [dev.slice(t << nv, (t + 1) << nv) for t in range(m)]
rr = random.Random(7)
pairs = [(rr.randrange(m), rr.randrange(m)) for _ in range(m)]
for _ in range(3):
    hal.execute(lambda ex: list(ex.bivariate_round_evals(sub, nv, pairs, 12345)))
