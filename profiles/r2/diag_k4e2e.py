"""K4 at a handful of points: device time of the graph-replayed launch sequence and e2e through BinnedEngine.evaluate."""
import json
import os
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench_workloads as wl                     # noqa: E402

os.chdir(tempfile.mkdtemp(prefix="bi_diag_"))
import torch                                     # noqa: E402
from blueice_b200.engine import BinnedEngine, MorphGrid   # noqa: E402

axes, edges, mus3, pmf, n_model, observed = wl.c3_arrays((200, 200, 20), 4, 3, (-1., 0., 1.), seed=3)
beng = BinnedEngine(MorphGrid(axes), mus3.reshape(27, 4), pmf, n_model, 0)
beng.set_observed(observed)
out = {}
for P in (1, 7, 256):
    zs3, mult3 = wl.scan_points(P, 3, 4, seed=31, z_range=(-1., 1.), mult_range=(0.8, 1.2))
    zs_d, mult_d, _, _, _ = beng._upload_points(zs3, mult3, None, None)
    zs_d, mult_d = zs_d.clone(), mult_d.clone()
    for _ in range(2):
        beng.run_device(P, zs_d, mult_d, None, None)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        beng.run_device(P, zs_d, mult_d, None, None)
    ms = []
    for k in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ref = beng.evaluate(zs3, mult3)
    for _ in range(4):
        got = beng.evaluate(zs3, mult3)
    assert np.array_equal(ref, got)
    t0 = time.perf_counter()
    n = 20 if P < 100 else 3
    for _ in range(n):
        beng.evaluate(zs3, mult3)
    out[P] = {"device_graph_ms": float(np.median(ms)), "e2e_ms": (time.perf_counter() - t0) / n * 1e3}
print("DIAG " + json.dumps(out))
