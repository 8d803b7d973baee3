"""2-GPU diagnostics: where the time of a point-sharded / event-sharded e2e call goes (torchrun, one process per GPU)."""
import cProfile
import json
import os
import pstats
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench_workloads as wl                                              # noqa: E402
from blueice_b200 import distributed as bdist                            # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
os.chdir(tempfile.mkdtemp(prefix="bi_diag2_"))
out = {}


def timed(fn, n=100, barrier=True):
    for _ in range(10):
        fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        if barrier:
            dist.barrier()
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    t = torch.tensor([float(np.median(ts)), float(np.min(ts)), float(np.percentile(ts, 90))], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) * 1e6 for x in t]


ll, d, names = wl.c2_api(2, 2, wl.ANCHORS_5, (100, 100), seed=1)
P = 4096
zs_all, mult_all = wl.scan_points(P * world, 2, 2, seed=2)
table_all = np.ascontiguousarray(np.column_stack([mult_all, zs_all]))
mine = table_all[rank * P:(rank + 1) * P]
sharded = bdist.PointShardedLikelihood(ll)
out["c2_weak_local_batch_us"] = timed(lambda: ll.batch(mine, names))
out["c2_weak_sharded_batch_us"] = timed(lambda: sharded.batch(table_all, names))
out["c2_weak_sharded_nobarrier_us"] = timed(lambda: sharded.batch(table_all, names), barrier=False)
half = table_all[:P]
out["c2_strong_local_batch_us"] = timed(lambda: ll.batch(half[rank * (P // world):(rank + 1) * (P // world)], names))
out["c2_strong_sharded_batch_us"] = timed(lambda: sharded.batch(half, names))
pg = next(iter(sharded._gathers.values()))
x = torch.zeros(pg.n, dtype=torch.float64, device=device)
for _ in range(20):
    pg.gather(x)
torch.cuda.synchronize()
dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200):
    pg.gather(x)
b.record()
torch.cuda.synchronize()
out["exchange_gather_%d_us" % pg.n] = a.elapsed_time(b) * 1e3 / 200
if rank == 0:
    pr = cProfile.Profile()
    pr.enable()
for _ in range(200):
    sharded.batch(table_all, names)
if rank == 0:
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(16)
    print("DIAG2 " + json.dumps(out))
dist.barrier()
dist.destroy_process_group()
