"""Steady-state efficiency of the DMMA K2 kernel: full units only, exactly `waves` units per resident warp.
    BI_MMA_TARGET_UNITS=<units> python profiles/steady_state.py [points] [superblocks_per_unit] [waves]
Prints the kernel time against the FP64-pipe ideal (8 DMMA-equivalent + 1 DMUL per point-event)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
points = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sb_per_unit = int(sys.argv[2]) if len(sys.argv) > 2 else 8
waves = int(sys.argv[3]) if len(sys.argv) > 3 else 1
groups = -(-points // 64)
units = 148 * 12 * waves
os.environ['BI_MMA_TARGET_UNITS'] = str(units)
import torch                                     # noqa: E402
from blueice_b200.engine import MorphGrid, UnbinnedEngine   # noqa: E402

n_events = (units // groups) * sb_per_unit * 512
axes = [np.linspace(-2, 2, 5), np.linspace(-2, 2, 5)]
rng = np.random.default_rng(0)
eng = UnbinnedEngine(MorphGrid(axes), rng.uniform(50, 100, (25, 2)))
eng.allocate_ps_anchor(n_events)
eng.ps_anchor.uniform_(1e-4, 1e-2)
zs = np.column_stack([rng.uniform(-0.9, -0.1, points), rng.uniform(0.1, 0.9, points)])     # one cell
mult = rng.uniform(0.5, 1.5, (points, 2))
zs_d, mult_d, _, _, _ = eng._upload_points(zs, mult, None, None)
zs_d, mult_d = zs_d.clone(), mult_d.clone()
outs = eng.run_fused(points, zs_d, mult_d, None, None)
_, v = eng.mma_workspace(points)


def plan_fn():
    eng.mma_plan(points, v, outs["status"])


def k2_fn():
    eng.mma_k2(v)


ms = []
for i in range(8):
    plan_fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    k2_fn()
    b.record()
    torch.cuda.synchronize()
    if i >= 3:
        ms.append(a.elapsed_time(b))
hdr = v["header"][:8].cpu().numpy()
t = float(np.mean(ms))
pe = float(points) * n_events
ideal_cycles = pe * (8 / 256 * 16 + 1 / 32 * 2) / (148 * 4)
print("points %d events %d units %d (groups %d x ranges %d, %d sb each): %.4f ms; pipe-ideal %.4f ms @1.92GHz -> %.1f%%; "
      "%.3e point-events/s; %.2f TFLOP/s (2CS flops)" % (points, n_events, hdr[3], hdr[0], hdr[1], hdr[2], t,
                                                          ideal_cycles / 1.92e6, 100 * ideal_cycles / 1.92e6 / t,
                                                          pe / t * 1e3, pe * 16 / t * 1e-9))
