"""Timing of the K-chunk DMMA kernel (k_unbinned_mma_wide) on scans with long contractions, next to the streaming kernel.

    python profiles/r2/wide_probe.py D S N P [stream]

Prints one JSON line: ms per evaluate() call (CUDA events around the engine call, mean of 5 after 2 warm-ups), the FP64
rate 2*K*P*N / t and the bits of agreement with the streaming kernel on the first 64 points."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from blueice_b200 import engine  # noqa: E402


def main():
    d, s, n, p = [int(v) for v in sys.argv[1:5]]
    with_stream = len(sys.argv) > 5
    rng = np.random.default_rng(1)
    axes = [np.sort(rng.uniform(-2, 2, 3)) for _ in range(d)]
    shape = [3] * d
    mus_anchor = rng.uniform(5, 500, shape + [s])
    ps_anchor = np.exp(rng.normal(-4, 2, shape + [s, n]))
    grid = engine.MorphGrid(axes)
    eng = engine.UnbinnedEngine(grid, mus_anchor.reshape(grid.n_anchors, s), 1e-12, None)
    eng.set_ps_anchor(ps_anchor)
    zs = np.column_stack([rng.uniform(a[0], a[-1], p) for a in axes]) if d else np.zeros((p, 0))
    mult = rng.uniform(0.5, 2, (p, s))
    out = {"D": d, "S": s, "K": (1 << d) * s, "N": n, "P": p, "wide_min_terms": os.environ.get("BI_MMA_WIDE_MIN_TERMS")}
    res = {}
    for mode in ([None, 'stream'] if with_stream else [None]):
        eng.force_kernel = mode
        pts = p if mode is None else min(p, 256)
        for _ in range(2):
            r = eng.evaluate(zs[:pts], mult[:pts])
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = eng.evaluate(zs[:pts], mult[:pts])
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[mode] = r
        ms = float(np.mean(ts))
        key = "mma" if mode is None else "stream"
        out[key + "_points"] = pts
        out[key + "_ms"] = ms
        out[key + "_tflops"] = 2.0 * out["K"] * pts * n / (ms * 1e-3) / 1e12
    if with_stream:
        k = min(64, len(res['stream']))
        out["max_abs_diff_vs_stream"] = float(np.max(np.abs(res[None][:k] - res['stream'][:k])))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
