"""Probe: torch symmetric memory (peer-mapped buffers over NVLink) on this box."""
import os
import time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n = 4096
t = symm_mem.empty(world * n, dtype=torch.float64, device=torch.device("cuda", rank))
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok; ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast", hdl.has_multicast_support, flush=True)
t.zero_()
hdl.barrier(channel=0)
src = torch.full((n,), float(rank + 1), dtype=torch.float64, device="cuda")
for r in range(world):
    peer = hdl.get_buffer(r, (world * n,), torch.float64)
    peer[rank * n:(rank + 1) * n].copy_(src)
hdl.barrier(channel=0)
torch.cuda.synchronize()
print(rank, "gathered", t.view(world, n)[:, 0].tolist(), flush=True)
# latency of barrier and of NCCL all_gather for comparison
out = torch.empty(world * n, dtype=torch.float64, device="cuda")
for name, fn in (("symm barrier", lambda: hdl.barrier(channel=0)),
                 ("nccl all_gather 32kB", lambda: dist.all_gather_into_tensor(out, src))):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100):
        fn()
    b.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(name, "%.1f us" % (a.elapsed_time(b) * 10), flush=True)
dist.destroy_process_group()
