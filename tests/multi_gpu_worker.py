"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun, NCCL).  Checks the three multi-GPU splits of
blueice_b200.distributed against single-GPU evaluations on the same inputs, bit for bit, and prints one line
`MULTI_GPU_OK {json}` from rank 0 when every rank passed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
        tests/multi_gpu_worker.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench_workloads as wl                                              # noqa: E402
from blueice_b200 import distributed as bdist                            # noqa: E402


def rank_ordered(rows):
    acc = rows[0].copy()
    for r in rows[1:]:
        acc = acc + r
    return acc


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.chdir(os.environ.get("TMPDIR", "/tmp"))
    report = {"world": world}
    from scipy import stats

    # ---- point sharding (config-2 shape, anchor-tensor engine): ragged point count, a rate prior, graph replays ----
    ll, d, names = wl.c2_api(2, 2, (-1., 0., 1.), (40, 30), n_events=5000, seed=11)
    ll.rate_parameters['bg'] = stats.norm(1, 0.2).logpdf
    zs, mult = wl.scan_points(301, 2, 2, seed=12, z_range=(-1., 1.))
    zs[7, 0] = 5.0                                                   # out of range -> -inf
    table = np.ascontiguousarray(np.column_stack([mult, zs]))
    single = ll.batch(table, names)
    sharded = bdist.PointShardedLikelihood(ll)
    for it in range(5):                                              # eager, capture, replays
        got = sharded.batch(table, names)
        assert np.array_equal(got, single), ("point sharding differs", it, np.abs(got - single).max())
    assert np.isneginf(single[7])
    pg = next(iter(sharded._gathers.values()))
    report["point_transport"] = "peer memory" if pg.fallback is None else "nccl fallback: " + pg.fallback
    assert pg.error() == 0
    # a second table with other values through the same graph
    table2 = table.copy()
    table2[:, 0] *= 1.01
    assert np.array_equal(sharded.batch(table2, names), ll.batch(table2, names))

    # equal shards, no priors: the gathered block is handed over without a host copy (rotating pinned landing buffers);
    # results that are still referenced must survive later calls, whatever the caller keeps (arrays, rows, slices)
    ll0, _, names0 = wl.c2_api(2, 2, (-1., 0., 1.), (40, 30), n_events=5000, seed=11)
    sharded0 = bdist.PointShardedLikelihood(ll0)
    tabs = [np.ascontiguousarray(table[:32 * world] * (1.0 + 0.001 * k)) for k in range(7)]
    for t in tabs:
        t[:, 2:] = table[:32 * world, 2:]                             # keep the shape parameters in range
    want = [ll0.batch(t, names0) for t in tabs]
    kept = []
    for k, t in enumerate(tabs):
        r = sharded0.batch(t, names0)
        assert np.array_equal(r, want[k]), ("equal shards differ", k)
        kept.append(r if k % 3 == 0 else (r[5:9] if k % 3 == 1 else None))   # whole array / a view / nothing
    for k, r in enumerate(kept):
        if r is not None:
            assert np.array_equal(r, want[k] if k % 3 == 0 else want[k][5:9]), ("a kept result was overwritten", k)
    report["landing_buffer_rotation"] = "ok"

    # ---- event sharding, anchor-tensor engine --------------------------------------------------------------------
    bounds = bdist.shard_bounds(len(d), world, align=512)
    pts = table[:23]
    parts = []
    for lo, hi in bounds:                                            # every shard alone on this GPU
        ll.set_data(d[lo:hi])
        parts.append(ll.batch_parts(pts, names))
    logsum_rows = [np.where(p[2] != 0, 0.0, p[0]) for p in parts]
    musum, status, priors = parts[0][1], parts[0][2], parts[0][3]
    expect = np.where(status != 0, -np.inf, priors + (-musum + rank_ordered(logsum_rows)))
    ll.set_data(d)
    whole = ll.batch(pts, names)
    ll.set_data(bdist.shard_events(d))
    ev = bdist.EventShardedLikelihood(ll)
    for it in range(5):
        got = ev.batch(pts, names)
        assert np.array_equal(got, expect), ("event sharding (anchor engine) differs", it)
    assert np.all(np.abs(got[np.isfinite(whole)] - whole[np.isfinite(whole)]) <= 1e-9 * len(d))
    assert np.array_equal(np.isfinite(got), np.isfinite(whole))
    one = ev(**dict(zip(names, [float(v) for v in pts[3]])))
    assert one == expect[3]

    # ---- event sharding, mixture engine (config-5 shape, small) ---------------------------------------------------
    ll5, _, names5 = wl.c2_api(6, 4, (-1., 0., 1.), (24, 20), n_events=1000, seed=5,
                               likelihood_config={'unbinned_engine': 'mixture'})
    lt = 60000.0 / float(np.sum(ll5.base_model.expected_events()))
    d5 = ll5.base_model.simulate_toys(1, livetime_days=lt, seed=50).to_records()
    rng = np.random.default_rng(51)
    x0 = np.concatenate([rng.uniform(0.8, 1.2, size=6), rng.uniform(-0.9, 0.9, size=4)])
    fd = np.repeat(x0[None, :], 11, 0)
    for j in range(10):
        fd[j + 1, j] += 1.4901161193847656e-08
    scan = np.column_stack([rng.uniform(0.8, 1.2, size=(20, 6)), rng.uniform(-0.9, 0.9, size=(20, 4))])
    bounds5 = bdist.shard_bounds(len(d5), world, align=512)
    for pts5 in (fd[:1], fd, scan):
        parts = []
        for lo, hi in bounds5:
            ll5.set_data(d5[lo:hi])
            parts.append(ll5.batch_parts(pts5, names5, livetime_days=lt))
        expect = np.where(parts[0][2] != 0, -np.inf,
                          parts[0][3] + (-parts[0][1] + rank_ordered([np.where(p[2] != 0, 0.0, p[0]) for p in parts])))
        ll5.set_data(d5)
        whole = ll5.batch(pts5, names5, livetime_days=lt)
        ll5.set_data(bdist.shard_events(d5))
        ev5 = bdist.EventShardedLikelihood(ll5)
        for it in range(5):
            got = ev5.batch(pts5, names5, livetime_days=lt)
            assert np.array_equal(got, expect), ("event sharding (mixture engine) differs", len(pts5), it)
        assert np.max(np.abs(got - whole)) <= 1e-9 * len(d5)
        rows = bdist.all_gather_rows(got)
        assert all(np.array_equal(rows[0], r) for r in rows), "ranks disagree"
    report["event_max_abs_diff_vs_one_gpu"] = float(np.max(np.abs(got - whole)))

    # ---- toy sharding (config-4 shape, small; ragged toy count) -----------------------------------------------------
    ll4, _, names4 = wl.c2_api(3, 3, (-1., 0., 1.), (40, 30), n_events=1000, seed=4)
    lt4 = 300.0 / float(np.sum(ll4.base_model.expected_events()))
    n_toys = 129
    zs4, mult4 = wl.scan_points(n_toys, 3, 3, seed=41, z_range=(-1., 1.))
    table4 = np.ascontiguousarray(np.column_stack([mult4, zs4]))
    ts4 = bdist.ToyShardedLikelihood(ll4)
    ts4.simulate(n_toys, livetime_days=lt4, seed=40)
    got4 = [ts4.batch_toys(table4, names4, livetime_days=lt4) for _ in range(5)]
    ll4.set_toy_data(ll4.base_model.simulate_toys(n_toys, livetime_days=lt4, seed=40))       # the same toys in one piece
    one4 = ll4.batch_toys(table4, names4, livetime_days=lt4)
    for g in got4:
        assert np.array_equal(g, one4), "sharded toys differ from the single-GPU toys"
    if ts4._gather is not None:
        assert ts4._gather.error() == 0

    ok = torch.ones(1, device="cuda")
    dist.all_reduce(ok)
    if rank == 0:
        assert int(ok.item()) == world
        print("MULTI_GPU_OK " + json.dumps(report), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
