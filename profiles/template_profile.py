"""ncu driver for the template-space kernels: one C5-shaped mixture evaluation (P = 1, N events) and one C4-shaped
toy sweep.   python profiles/template_profile.py [n_events] [n_toys]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_workloads as wl                                              # noqa: E402
from template_bench import build, draw_events                            # noqa: E402

n_events = int(sys.argv[1]) if len(sys.argv) > 1 else 100000000
n_toys = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
dev = torch.device("cuda:0")
eng, tb, mb, edges = build(6, 4, False, 'mixture')
eng.set_datasets(draw_events(tb, mb, edges, n_events, dev, 5))
rng = np.random.default_rng(5)
z0, m0 = rng.uniform(-1.9, 1.9, size=(1, 4)), rng.uniform(0.8, 1.2, size=(1, 6))
for _ in range(3):
    print("C5 mixture P=1:", eng.evaluate(z0, m0))
del eng
torch.cuda.empty_cache()
eng, tb, mb, edges = build(3, 3, False, 'exact')
sizes = np.random.default_rng(4).poisson(1000, size=n_toys)
offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
eng.set_datasets(draw_events(tb, mb, edges, int(offsets[-1]), dev, 1), offsets)
zs, mult = wl.scan_points(n_toys, 3, 3, seed=4)
for _ in range(2):
    print("C4 toys:", eng.evaluate_toys(zs, mult)[:3])
