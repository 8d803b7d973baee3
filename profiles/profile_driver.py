"""Minimal launch sequence for ncu: config-2 scan through ll.batch (grouped K2) and one P=1 evaluation of
an 8M-event tensor (streaming K2).  Usage (on the GPU box):
    python profiles/profile_driver.py && ncu --set full --clock-control none --import-source on \
        -k regex:k_unbinned -c 4 -o gpurun_out/prof python profiles/profile_driver.py
"""
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench_workloads as wl                     # noqa: E402

os.chdir(tempfile.mkdtemp(prefix="bi_prof_"))
import torch                                     # noqa: E402
from blueice_b200.engine import MorphGrid, UnbinnedEngine   # noqa: E402

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ll, d, names = wl.c2_api(2, 2, wl.ANCHORS_5, (100, 100), seed=1)
zs, mult = wl.scan_points(4096, 2, 2, seed=2)
table = np.column_stack([mult, zs])
for _ in range(n_rep):
    out = ll.batch(table, names)
print("grouped: logL[0..2] =", out[:3])

eng = ll._engine
eng.force_kernel = os.environ.get('BI_KERNEL') or None
n_big = 8 * 1024 * 1024
big = UnbinnedEngine(MorphGrid(eng.grid.axes), eng.mus_anchor_host)
big.allocate_ps_anchor(n_big)
n = eng.n_events
for r in range(-(-n_big // n)):
    lo, hi = r * n, min((r + 1) * n, n_big)
    big.ps_anchor[:, :, lo:hi].copy_(eng.ps_anchor[:, :, :hi - lo])
big.force_kernel = os.environ.get('BI_STREAM_KERNEL') or None
for _ in range(n_rep):
    one = big.evaluate(zs[:1], mult[:1])
print("P=1 over 8 Mi events: logL =", one)
torch.cuda.synchronize()
