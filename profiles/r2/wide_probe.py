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
    d, s, n = [int(v) for v in sys.argv[1:4]]
    p_list = [int(v) for v in sys.argv[4].split(',')]
    p = max(p_list)
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
    if len(p_list) > 1:                      # a list of batch sizes: K2 alone at every size, one JSON line each
        for pp in p_list:
            o2 = dict(D=d, S=s, K=out["K"], N=n, P=pp, wide_min_terms=out["wide_min_terms"],
                      wide_min_points=os.environ.get("BI_MMA_WIDE_MIN_POINTS"))
            o2.update(k2_alone(eng, zs[:pp], mult[:pp], n, out["K"]))
            print(json.dumps(o2))
        return
    out.update(k2_alone(eng, zs, mult, n, out["K"]))
    if with_stream:
        k = min(64, len(res['stream']))
        out["max_abs_diff_vs_stream"] = float(np.max(np.abs(res[None][:k] - res['stream'][:k])))
    print(json.dumps(out))


def k2_alone(eng, zs, mult, n, K):
    """K2 alone (coefficient packing included where the K-chunk kernel runs) on the workspace of a fused call."""
    p = len(zs)
    out = {}
    eng.force_kernel = None
    z_d, m_d, _, _, _ = eng._upload_points(zs, mult, None, None)
    z_d, m_d = z_d.clone(), m_d.clone()
    outs = eng.run_fused(p, z_d, m_d, None, None)
    _, v = eng.mma_workspace(p)
    ts = []
    for _ in range(7):
        eng.mma_plan(p, v, outs["status"])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.mma_k2(v)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    out["k2_ms"] = float(np.mean(ts[2:]))
    out["k2_tflops"] = 2.0 * K * p * n / (out["k2_ms"] * 1e-3) / 1e12
    out["k2_hbm_gbs_if_one_pass"] = 8.0 * K * n / (out["k2_ms"] * 1e-3) / 1e9
    return out


if __name__ == "__main__":
    main()
