"""ncu driver: config-5-shaped mixture evaluation at P = 11 (grouped K5b on the FP64 tensor pipe).
    python profiles/r2/prof_k5b.py [n_events] [P]"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from template_bench import build, draw_events                            # noqa: E402

n_events = int(sys.argv[1]) if len(sys.argv) > 1 else 100000000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 11
dev = torch.device("cuda:0")
eng, tb, mb, edges = build(6, 4, False, 'mixture')
eng.set_datasets(draw_events(tb, mb, edges, n_events, dev, 5))
rng = np.random.default_rng(5)
z0, m0 = rng.uniform(-1.9, 1.9, size=(P, 4)), rng.uniform(0.8, 1.2, size=(P, 6))
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = eng.evaluate(z0, m0)
    b.record()
    torch.cuda.synchronize()
    print("C5 mixture P=%d: %.4f ms" % (P, a.elapsed_time(b)), r[:2])
