"""N-GPU timeline of the point-sharded e2e call: per rank and call the host clock (CLOCK_MONOTONIC: one host, comparable
across the ranks) at the start and at the end, with and without a barrier in front of every call, next to the local
ll.batch under the same load.  torchrun, one process per GPU; writes gpurun_out/diagN_rank<r>.json."""
import json
import os
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
outdir = os.path.join(REPO, "gpurun_out")
os.chdir(tempfile.mkdtemp(prefix="bi_diagn_"))
flush = torch.empty(512 << 20, dtype=torch.uint8, device=device)
out = {"rank": rank, "world": world}


def timeline(fn, n, barrier, flush_l2=True):
    for _ in range(10):
        fn()
    t0s, t1s = [], []
    for _ in range(n):
        if flush_l2:
            flush.fill_(1)
        torch.cuda.synchronize()
        if barrier == "nccl":
            dist.barrier()
        elif barrier == "pad":
            pg.barrier()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        t1s.append(time.perf_counter())
        t0s.append(t0)
    return {"t0": t0s, "t1": t1s}


ll, d, names = wl.c2_api(2, 2, wl.ANCHORS_5, (100, 100), seed=1)
P = 4096
zs_all, mult_all = wl.scan_points(P * world, 2, 2, seed=2)
table_all = np.ascontiguousarray(np.column_stack([mult_all, zs_all]))
mine = table_all[rank * P:(rank + 1) * P]
sharded = bdist.PointShardedLikelihood(ll)
sharded.batch(table_all, names)
pg = next(iter(sharded._gathers.values()))
N = 100
out["local_nccl"] = timeline(lambda: ll.batch(mine, names), N, "nccl")
out["sharded_nccl"] = timeline(lambda: sharded.batch(table_all, names), N, "nccl")
out["sharded_pad"] = timeline(lambda: sharded.batch(table_all, names), N, "pad")
out["sharded_free"] = timeline(lambda: sharded.batch(table_all, names), N, None)
out["sharded_free_noflush"] = timeline(lambda: sharded.batch(table_all, names), N, None, flush_l2=False)
# the exchange alone, back to back on the device
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
out["exchange_us"] = a.elapsed_time(b) * 1e3 / 200
json.dump(out, open(os.path.join(outdir, "diag%d_rank%d.json" % (world, rank)), "w"))
dist.barrier()
dist.destroy_process_group()
