"""One GPU: where an ll.batch call of the config-5 mixture engine spends its time when the shard is small
(1.25e7 events = one rank's share at N = 8): e2e wall time, the device span, a cProfile of the host side.
    python profiles/r2/prof_c5_host.py [n_events]"""
import cProfile
import os
import pstats
import sys
import tempfile
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench_workloads as wl                                              # noqa: E402

n_target = int(float(sys.argv[1])) if len(sys.argv) > 1 else 12500000
os.chdir(tempfile.mkdtemp(prefix="bi_c5host_"))
ll, _, names = wl.c2_api(6, 4, wl.ANCHORS_5, (100, 100), n_events=1000, seed=5, likelihood_config={'unbinned_engine': 'mixture'})
base_mu = float(np.sum(ll.base_model.expected_events()))
lt = n_target / base_mu
td = ll.base_model.simulate_toys(1, livetime_days=lt, seed=50)
d = np.zeros(td.n_events, dtype=[('source', int)] + [(name, float) for name in td.dims])
host = td.coords.cpu().numpy()
for k, name in enumerate(td.dims):
    d[name] = host[k]
ll.set_data(d)
rng = np.random.default_rng(51)
x0 = np.concatenate([rng.uniform(0.8, 1.2, size=6), rng.uniform(-1.9, 1.9, size=4)])
fd = np.repeat(x0[None, :], 11, 0)
for j in range(10):
    fd[j + 1, j] += 1.4901161193847656e-08
scan = np.column_stack([rng.uniform(0.8, 1.2, size=(64, 6)), rng.uniform(-1.9, 1.9, size=(64, 4))])
for P, table in ((1, fd[:1]), (11, fd), (64, scan)):
    for fn_name, fn in (("batch", lambda: ll.batch(table, names, livetime_days=lt)),
                        ("batch_parts", lambda: ll.batch_parts(table, names, livetime_days=lt))):
        for _ in range(10):
            fn()
        ts, ds = [], []
        for _ in range(50):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record()
            fn()
            b.record()
            ts.append(time.perf_counter() - t0)
            torch.cuda.synchronize()
            ds.append(a.elapsed_time(b))
        print("C5HOST N=%d P=%d %s: e2e %.1f us, event span %.1f us" % (td.n_events, P, fn_name, np.median(ts) * 1e6, np.median(ds) * 1e3))
table = fd
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    ll.batch_parts(table, names, livetime_days=lt)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
