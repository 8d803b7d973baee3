"""Round-2 diagnostics (1 GPU): where the time of an e2e call goes at small batches, and the K4 / K5 baselines.
    python profiles/r2/diag1.py [case ...]      cases: c2 c5 k4 k5 (default: all)"""
import cProfile
import json
import os
import pstats
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench_workloads as wl                     # noqa: E402

os.chdir(tempfile.mkdtemp(prefix="bi_diag_"))
import torch                                     # noqa: E402

cases = sys.argv[1:] or ["c2", "c5", "k4", "k5"]
out = {}


def timeit(fn, n, sync=True):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    if sync:
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


def dev_time(fn, n):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


if "c1" in cases:
    ll1, d1, names1 = wl.c1_api(seed=0)
    kw1 = {'s0_rate_multiplier': 1.1, 'mu': 0.3}
    lat = timeit(lambda: ll1(**kw1), 2000)
    pts = np.array([[1.1, 0.3]])
    lat_b = timeit(lambda: ll1.batch(pts, names1), 1000)
    eng = ll1._engine
    pin, run = eng.scalar_runner(False)
    pin[:] = [0.3, 1.1]
    lat_r = timeit(run, 2000)
    out["c1_single_call"] = {"ll_kwargs_us": lat * 1e6, "ll_batch_P1_us": lat_b * 1e6, "scalar_runner_us": lat_r * 1e6,
                             "small_path": bool(eng._small_ok(1)), "n_events": int(len(d1))}
    rng1 = np.random.default_rng(1)
    pts1 = np.column_stack([rng1.uniform(0.5, 2.0, 4096), rng1.uniform(-2.0, 2.0, 4096)])
    firsts = []
    for _ in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ll1.batch(pts1, names1)
        firsts.append((time.perf_counter() - t0) * 1e3)
    out["c1_batch4096_first_calls_ms"] = firsts
    out["c1_batch4096_steady_ms"] = timeit(lambda: ll1.batch(pts1, names1), 50) * 1e3
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(2000):
        ll1(**kw1)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(12)

if "c2" in cases:
    ll, d, names = wl.c2_api(2, 2, wl.ANCHORS_5, (100, 100), seed=1)
    eng = ll._engine
    zs, mult = wl.scan_points(4096, 2, 2, seed=2)
    table = np.ascontiguousarray(np.column_stack([mult, zs]))
    res = {}
    for P in (1, 64, 512, 1024, 2048, 4096):
        t = table[:P]
        wall = timeit(lambda: ll.batch(t, names), 100)
        zs_d, mult_d, _, _, _ = eng._upload_points(zs[:P], mult[:P], None, None)
        zs_d, mult_d = zs_d.clone(), mult_d.clone()
        dev = dev_time(lambda: eng.run_fused(P, zs_d, mult_d, None, None), 20)
        res[P] = {"wall_us": wall * 1e6, "device_us": dev * 1e3}
    out["c2_batch"] = res
    pr = cProfile.Profile()
    t = table[:512]
    pr.enable()
    for _ in range(300):
        ll.batch(t, names)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)

if "c5" in cases:
    ll, _, names = wl.c2_api(6, 4, wl.ANCHORS_5, (100, 100), n_events=1000, seed=5,
                             likelihood_config={'unbinned_engine': 'mixture'})
    base_mu = float(np.sum(ll.base_model.expected_events()))
    n_target = 12500000
    lt = n_target / base_mu
    d = ll.base_model.simulate_toys(1, livetime_days=lt, seed=50).to_records()
    ll.set_data(d)
    rng = np.random.default_rng(51)
    x0 = np.concatenate([rng.uniform(0.8, 1.2, size=6), rng.uniform(-1.9, 1.9, size=4)])
    fd = np.repeat(x0[None, :], 11, 0)
    for j in range(10):
        fd[j + 1, j] += 1.4901161193847656e-08
    eng = ll._engine
    res = {}
    for P in (1, 11):
        t = fd[:P]
        wall = timeit(lambda: ll.batch(t, names, livetime_days=lt), 100)
        zs, mult = ll._rows_from_params(t, names)
        sched, _ = eng.single_schedule(zs)
        zs_d, mult_d, scale_d, _, _ = eng._upload_points(zs, mult, np.full(P, lt), None)
        dev = dev_time(lambda: eng.run_one_call(P, sched, zs_d, mult_d, scale_d, None), 20)
        res[P] = {"wall_us": wall * 1e6, "device_us": dev * 1e3}
    out["c5_1.25e7_events"] = res
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        ll.batch(fd, names, livetime_days=lt)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
    del ll, eng, d
    torch.cuda.empty_cache()

if "k4" in cases:
    from blueice_b200.engine import BinnedEngine, MorphGrid
    axes, edges, mus3, pmf, n_model, observed = wl.c3_arrays((200, 200, 20), 4, 3, (-1., 0., 1.), seed=3)
    beng = BinnedEngine(MorphGrid(axes), mus3.reshape(27, 4), pmf, n_model, 0)
    beng.set_observed(observed)
    res = {}
    for P in (1, 7, 16, 256):
        zs3, mult3 = wl.scan_points(P, 3, 4, seed=31, z_range=(-1., 1.), mult_range=(0.8, 1.2))
        zs_d, mult_d, _, _, _ = beng._upload_points(zs3, mult3, None, None)
        zs_d, mult_d = zs_d.clone(), mult_d.clone()
        res[P] = {"device_ms": dev_time(lambda: beng.run_device(P, zs_d, mult_d, None, None), 6)}
    out["k4_config3"] = res
    del beng
    torch.cuda.empty_cache()

if "k4big" in cases or "k4p1" in cases:
    # one size only, few launches: for ncu captures of the P = 256 scan / the P = 1 evaluation
    from blueice_b200.engine import BinnedEngine, MorphGrid
    axes, edges, mus3, pmf, n_model, observed = wl.c3_arrays((200, 200, 20), 4, 3, (-1., 0., 1.), seed=3)
    beng = BinnedEngine(MorphGrid(axes), mus3.reshape(27, 4), pmf, n_model, 0)
    beng.set_observed(observed)
    P = 256 if "k4big" in cases else 1
    zs3, mult3 = wl.scan_points(P, 3, 4, seed=31, z_range=(-1., 1.), mult_range=(0.8, 1.2))
    zs_d, mult_d, _, _, _ = beng._upload_points(zs3, mult3, None, None)
    zs_d, mult_d = zs_d.clone(), mult_d.clone()
    out["k4_P%d" % P] = {"device_ms": dev_time(lambda: beng.run_device(P, zs_d, mult_d, None, None), 3)}
    del beng
    torch.cuda.empty_cache()

if "k5" in cases:
    ll, _, names = wl.c2_api(3, 3, wl.ANCHORS_5, (100, 100), n_events=1000, seed=4)
    lt = 1000.0 / float(np.sum(ll.base_model.expected_events()))
    T = 100000
    toys = ll.base_model.simulate_toys(T, livetime_days=lt, seed=40)
    ll.set_toy_data(toys)
    zs, mult = wl.scan_points(T, 3, 3, seed=41)
    eng = ll._toy_engine
    zs_d, mult_d, scale_d, _, _ = eng._upload_points(zs, mult, np.full(T, lt), None)
    sched = eng.toy_schedule()
    out["k5_toys_1e5"] = {"device_ms": dev_time(lambda: eng.run_one_call(T, sched, zs_d, mult_d, scale_d, None), 6),
                          "events": int(toys.n_events), "bin_major": sched.get("bm") is not None,
                          "n_tasks": None if sched.get("bm") is None else sched["bm"]["n_tasks"]}
    if sched.get("bm") is not None:                    # the same sweep toy by toy (K5): bit-identical, slower
        _, logl, _ = eng.run_one_call(T, sched, zs_d, mult_d, scale_d, None)
        a = logl[:T].cpu().numpy().copy()
        plain = dict(sched)
        plain["bm"] = None
        out["k5_toys_1e5"]["toy_major_device_ms"] = dev_time(lambda: eng.run_one_call(T, plain, zs_d, mult_d, scale_d, None), 4)
        _, logl, _ = eng.run_one_call(T, plain, zs_d, mult_d, scale_d, None)
        b = logl[:T].cpu().numpy().copy()
        out["k5_toys_1e5"]["bit_identical"] = bool(np.array_equal(a, b))
        out["k5_toys_1e5"]["finite"] = int(np.isfinite(a).sum())

print("DIAG " + json.dumps(out))
