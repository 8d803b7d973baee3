"""Kernel-only timing of the grouped K5b kernel (k_mixture_partials_mma) on a config-5-shaped dataset:
    python profiles/r2/prof_k5b_kernel.py [n_events] [P ...]
prints, per P, the whole device sequence (K1 + mix + K5b + finalize) and the streaming kernel alone (mean of 10)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from template_bench import build, draw_events                            # noqa: E402

n_events = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
Ps = [int(v) for v in sys.argv[2:]] or [11, 64]
dev = torch.device("cuda:0")
eng, tb, mb, edges = build(6, 4, False, 'mixture')
eng.set_datasets(draw_events(tb, mb, edges, n_events, dev, 5))
rng = np.random.default_rng(5)
for P in Ps:
    zs, mult = rng.uniform(-1.9, 1.9, size=(P, 4)), rng.uniform(0.8, 1.2, size=(P, 6))
    if P == 11:                                                   # a finite-difference batch: one cell
        zs[:] = zs[0]
    sched, _ = eng.single_schedule(zs)
    zs_d, mult_d, scale_d, _, _ = eng._upload_points(zs, mult, None, None)
    dm, km = [], []
    for k in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.run_one_call(P, sched, zs_d, mult_d, scale_d, None)
        b.record()
        torch.cuda.synchronize()
        if k > 1:
            dm.append(a.elapsed_time(b))
    o = eng._setup_terms(P, zs_d, mult_d, scale_d, None)
    eng.run_schedule(sched, o)
    for k in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.mixture_kernel_only(sched, o)
        b.record()
        torch.cuda.synchronize()
        if k > 1:
            km.append(a.elapsed_time(b))
    print("K5B N=%d P=%d groups=%d gp=%d: device %.4f ms, kernel %.4f ms (min %.4f)"
          % (n_events, P, sched["n_groups"], sched["group_points"], np.mean(dm), np.mean(km), np.min(km)))
