"""Where the time of inference.bestfit_toys goes (host profile; config-4-shaped model, 5000 toys)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_workloads as wl                                              # noqa: E402
from blueice_b200.inference import bestfit_toys                           # noqa: E402

os.chdir("/tmp")
ll, _, names = wl.c2_api(3, 3, wl.ANCHORS_5, (100, 100), n_events=1000, seed=4)
lt = 1000.0 / float(np.sum(ll.base_model.expected_events()))
ll.set_toy_data(ll.base_model.simulate_toys(5000, livetime_days=lt, seed=40))
bestfit_toys(ll, livetime_days=lt, max_iter=3)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
_, _, info = bestfit_toys(ll, livetime_days=lt)
pr.disable()
print("fit: %.3f s, %d iterations, %d evaluations" % (time.perf_counter() - t0, info["iterations"], info["evaluations"]))
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
