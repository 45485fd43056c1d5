"""Run under torchrun on N GPUs: sharded bivariate sumcheck (binius_b200/sharding.py) on real B200s
with NCCL for the combine, checked round-by-round against the CPU oracle.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/sharded_sumcheck_check.py
"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import binius_b200
from binius_b200 import sharding
from oracle import binding as orc

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hal = binius_b200.B200Layer(local)
n_vars, m = 14, 6
rng = random.Random(1)
mls = [orc.rand_b128(100 + t, 1 << n_vars) for t in range(m)]
pairs = [(0, 1), (2, 3), (4, 5), (1, 4)]
alphas = [rng.getrandbits(128) for _ in range(n_vars)]
chs = [rng.getrandbits(128) for _ in range(n_vars)]
sc = sharding.ShardedBivariateSumcheck(hal, mls, n_vars, pairs, world, rank, dist, comm_device=f"cuda:{local}")
cur = [x.copy() for x in mls]
for r in range(n_vars):
    exp = tuple(orc.bivariate_round_evals(cur, n_vars - r, pairs, alphas[r]))
    got = sc.round_evals(alphas[r])
    assert got == exp, (rank, r)
    sc.fold(chs[r])
    cur = [orc.extrapolate_line(x[: len(x) // 2], x[len(x) // 2:], chs[r]) for x in cur]
assert sc.finish() == [orc.to_ints(x)[0] for x in cur]
dist.barrier()
if rank == 0:
    print(f"sharded sumcheck ok on {world} GPUs: {n_vars} rounds bit-exact vs oracle")
dist.destroy_process_group()
