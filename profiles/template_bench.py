"""Throughput of the template-space kernel (K5) at BASELINE configs 4 and 5 shapes (profiling aid, not bench.py).

    python profiles/template_bench.py [c4_toys] [c5_events]

config 4: 3 sources, 3 shape parameters x 5 anchors (125 anchors), 100x100 templates, T toys x ~1000 events,
          one parameter point per toy.   config 5: 6 sources, 4 shape parameters x 5 anchors (625 anchors),
          N events, P = 1 and P = 11 (a finite-difference batch around one point).
Events are drawn on the device with torch (plumbing); timing = CUDA events around K1 + K5 + finalize."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_workloads as wl                                              # noqa: E402
from blueice_b200.engine import MorphGrid, TemplateUnbinnedEngine        # noqa: E402


def draw_events(templates_base, mus_base, edges, n, device, seed):
    """n events from the base mixture: bin ~ mixture pmf, uniform inside the bin (device-side)."""
    g = torch.Generator(device=device).manual_seed(seed)
    vol = np.outer(np.diff(edges[0]), np.diff(edges[1]))
    pmf = sum(m * t * vol for m, t in zip(mus_base, templates_base)).ravel()
    pmf = torch.from_numpy(pmf / pmf.sum()).to(device)
    cdf = torch.cumsum(pmf, 0)
    u = torch.rand(n, generator=g, device=device, dtype=torch.float64)
    flat = torch.searchsorted(cdf, u).clamp_(max=len(pmf) - 1)
    ny = len(edges[1]) - 1
    ix, iy = flat // ny, flat % ny
    e0 = torch.from_numpy(edges[0]).to(device)
    e1 = torch.from_numpy(edges[1]).to(device)
    x = e0[ix] + torch.rand(n, generator=g, device=device, dtype=torch.float64) * (e0[ix + 1] - e0[ix])
    y = e1[iy] + torch.rand(n, generator=g, device=device, dtype=torch.float64) * (e1[iy + 1] - e1[iy])
    return torch.stack([x, y])


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def build(n_sources, n_shape, bin_major, mode='exact'):
    axes, edges, templates, mus = wl.c2_arrays(n_sources, n_shape, wl.ANCHORS_5, (100, 100))
    grid = MorphGrid(axes)
    rows = templates.reshape((grid.n_anchors * n_sources, 100, 100))
    eng = TemplateUnbinnedEngine(grid, mus.reshape(grid.n_anchors, n_sources), rows, edges, 'linear', mode=mode)
    centre = tuple(len(a) // 2 for a in axes)
    return eng, templates[centre], mus[centre], edges


def device_eval(eng, sched, zs, mult):
    P = len(mult)
    zs_d, mult_d, scale_d, eff_d, _ = eng._upload_points(zs, mult, None, None)

    def run():
        o = eng._setup_terms(P, zs_d, mult_d, scale_d, eff_d)
        eng.run_schedule(sched, o)
    return run


def main():
    n_toys = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    n_c5 = int(sys.argv[2]) if len(sys.argv) > 2 else 20000000
    dev = torch.device("cuda:0")
    for bin_major in ((False,) if n_toys else ()):
        eng, tb, mb, edges = build(3, 3, bin_major)
        rng = np.random.default_rng(4)
        sizes = rng.poisson(1000, size=n_toys)
        offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        coords = draw_events(tb, mb, edges, int(offsets[-1]), dev, 1)
        t0 = time.perf_counter()
        eng.set_datasets(coords, offsets)
        torch.cuda.synchronize()
        t_prep = time.perf_counter() - t0
        zs, mult = wl.scan_points(n_toys, 3, 3, seed=4)
        ms = timed(device_eval(eng, eng.toy_schedule(), zs, mult))
        t0 = time.perf_counter()
        res = eng.evaluate_toys(zs, mult)
        t_e2e = time.perf_counter() - t0
        n_ev = int(offsets[-1])
        print("C4 bin_major=%d: %d toys, %d events, K=%d: device %.2f ms -> %.3e point-events/s, %.3e toys/s; "
              "e2e %.1f ms; prepare %.1f ms; finite=%d" % (bin_major, n_toys, n_ev, eng.n_terms, ms, n_ev / ms * 1e3,
                                                           n_toys / ms * 1e3, t_e2e * 1e3, t_prep * 1e3,
                                                           int(np.isfinite(res).sum())), flush=True)
        del eng, coords
        torch.cuda.empty_cache()
    modes = ((False, 'mixture'), (False, 'exact')) if os.environ.get('TPL_EXACT') else ((False, 'mixture'),)
    for bin_major, mode in (modes if n_c5 else ()):
        eng, tb, mb, edges = build(6, 4, bin_major, mode)
        coords = draw_events(tb, mb, edges, n_c5, dev, 5)
        eng.set_datasets(coords)
        if os.environ.get('TPL_WIDE_MIN'):                                # K5b: dataset size from which groups hold 16 points
            eng.mix_wide_min_superblocks = int(os.environ['TPL_WIDE_MIN'])
        torch.cuda.synchronize()
        rng = np.random.default_rng(5)
        z0 = rng.uniform(-1.9, 1.9, size=(1, 4))
        m0 = rng.uniform(0.8, 1.2, size=(1, 6))
        for P in tuple(int(v) for v in os.environ.get('TPL_POINTS', '1,11').split(',')):
            zs = np.repeat(z0, P, 0)
            mult = np.repeat(m0, P, 0)
            for j in range(1, P):                                     # forward-difference batch: one parameter each
                k = (j - 1) % 10
                if k < 6:
                    mult[j, k] += 1.5e-8 * (1 + (j - 1) // 10)
                else:
                    zs[j, k - 6] += 1.5e-8 * (1 + (j - 1) // 10)
            sched, order = eng.single_schedule(zs)
            ms = timed(device_eval(eng, sched, zs, mult))
            bytes_ev = 4 + 8 * eng.n_space
            print("C5 %s bin_major=%d: N=%d, P=%d, K=%d, groups=%d: device %.3f ms -> %.3e point-events/s; "
                  "prepared-event stream %.0f GB/s per group pass"
                  % (mode, bin_major, n_c5, P, eng.n_terms, sched["n_groups"], ms, P * n_c5 / ms * 1e3,
                     sched["n_groups"] * n_c5 * bytes_ev / ms / 1e6), flush=True)
        del eng, coords
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
