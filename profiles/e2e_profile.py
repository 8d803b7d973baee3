"""Host-side profile of the e2e path (ll.batch on the config-2 scan): where the time outside the kernels goes."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_workloads as wl                                              # noqa: E402

os.chdir("/tmp")
ll, d, names = wl.c2_api(2, 2, wl.ANCHORS_5, (100, 100), seed=1)
zs, mult = wl.scan_points(4096, 2, 2, seed=2)
table = np.ascontiguousarray(np.column_stack([mult, zs]))
for _ in range(20):
    ll.batch(table, names)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    ll.batch(table, names)
print("ll.batch P=4096: %.1f us per call" % ((time.perf_counter() - t0) / 200 * 1e6))
one = table[:1]
for _ in range(20):
    ll.batch(one, names)
t0 = time.perf_counter()
for _ in range(500):
    ll.batch(one, names)
print("ll.batch P=1: %.1f us per call" % ((time.perf_counter() - t0) / 500 * 1e6))
kw = dict(zip(names, [float(v) for v in one[0]]))
t0 = time.perf_counter()
for _ in range(500):
    ll(**kw)
print("ll(**kw): %.1f us per call" % ((time.perf_counter() - t0) / 500 * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    ll.batch(table, names)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
