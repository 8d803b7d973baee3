"""Full-size parity (BASELINE.json configs 2 and 3) through size-independent properties plus a few
direct oracle evaluations (the oracle needs ~5 ms per config-2 point and ~1 s per config-3 point)."""
import numpy as np
import pytest

import bench_workloads as wl
from oracle import hist as ohist
from oracle.pipeline import BinnedOracle, UnbinnedOracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def config2():
    import os
    import tempfile
    os.chdir(tempfile.mkdtemp(prefix="bi_full_"))
    ll, d, names = wl.c2_api(2, 2, wl.ANCHORS_5, (100, 100), seed=1)
    zs, mult = wl.scan_points(4096, 2, 2, seed=2)
    return ll, d, names, zs, mult


def test_config2_scan_properties(config2):
    ll, d, names, zs, mult = config2
    n = len(d)
    assert 99000 < n < 101000
    table = np.column_stack([mult, zs])
    eng = ll._engine
    res = {}
    for mode in ('stream', None):
        eng.force_kernel = mode
        res[mode] = ll.batch(table, names)
    eng.force_kernel = None
    # the two kernels agree to the stated tolerance; batch composition and scalar calls: identical bits
    assert eng.plan(zs).kernel == 'mma'                  # the default path is the DMMA kernel
    dm = np.abs(res[None] - res['stream'])
    assert np.all(dm <= 2e-13 * (np.abs(res['stream']) + n)), dm.max()
    for i in (0, 1, 1000, 4095):
        assert ll(**dict(zip(names, [float(v) for v in table[i]]))) == res[None][i]
    perm = np.random.default_rng(0).permutation(4096)[:777]
    assert np.array_equal(ll.batch(table[perm], names), res[None][perm])
    # direct comparison with the oracle on ALL 4096 points of the scan (~2 ms per point on one core)
    axes, edges, templates, mus = wl.c2_arrays(2, 2, wl.ANCHORS_5, (100, 100))
    orc = UnbinnedOracle(axes, mus).set_data_from_templates(templates, edges, [d['cs1'], d['cs2']])
    ref = orc.batch(zs, mult)
    diff = np.abs(res[None] - ref)
    assert np.all(diff <= 1e-9 * n), diff.max()
    assert np.all(diff <= 2e-13 * (np.abs(ref) + n)), diff.max()
    # the K3 gather reproduces the reference's per-event anchor tensor bit for bit (two full rows)
    dev = eng.ps_anchor[:, :, :n].cpu().numpy().reshape(5, 5, 2, n)
    for (a, b, s) in ((0, 4, 1), (3, 2, 0)):
        assert np.array_equal(dev[a, b, s], ohist.lookup_linear(templates[a, b, s], edges, [d['cs1'], d['cs2']]))
    # out-of-range and unphysical points inside a big batch
    table2 = table.copy()
    table2[5, 2] = 2.5
    table2[6, 0] = -0.1
    out = ll.batch(table2, names)
    assert out[5] == -np.inf and out[6] == -np.inf
    keep = np.ones(4096, dtype=bool)
    keep[[5, 6]] = False
    assert np.array_equal(out[keep], res[None][keep])


def test_config2_event_shards_add_up(config2):
    """Event sharding: superblock-aligned shards, per-shard log sums add up to the total."""
    from blueice_b200 import distributed as bdist
    from blueice_b200.engine import UnbinnedEngine
    ll, d, names, zs, mult = config2
    eng = ll._engine
    n = len(d)
    total, musum, status = eng.evaluate(zs[:300], mult[:300], return_parts=True)
    bounds = bdist.shard_bounds(n, 4, align=512)
    acc = np.zeros(300)
    for lo, hi in bounds:
        shard = UnbinnedEngine(eng.grid, eng.mus_anchor_host)
        shard.allocate_ps_anchor(hi - lo)
        shard.ps_anchor[:, :, :hi - lo].copy_(eng.ps_anchor[:, :, lo:hi])
        part, musum_s, _ = shard.evaluate(zs[:300], mult[:300], return_parts=True)
        assert np.array_equal(musum_s, musum)
        acc = acc + part
    np.testing.assert_allclose(acc, total, rtol=2e-14)
    assert np.all(np.abs(acc - total) <= 1e-9 * n)


def test_config2_anchor_points_and_idempotence(config2):
    ll, d, names, zs, mult = config2
    base = ll.batch(np.column_stack([mult[:64], zs[:64]]), names)
    ll.set_data(d)                                      # rebuilding the device tensor changes nothing
    assert np.array_equal(ll.batch(np.column_stack([mult[:64], zs[:64]]), names), base)
    # on anchors the morph weights are exactly {0, 1}: the value equals the anchor model's own likelihood
    from blueice_b200.engine import MorphGrid, UnbinnedEngine
    eng = ll._engine
    n = len(d)
    for (a, b) in ((0, 0), (2, 2), (4, 1), (4, 4)):
        z = np.array([[wl.ANCHORS_5[a], wl.ANCHORS_5[b]]])
        got = eng.evaluate(z, mult[:1])
        flat = UnbinnedEngine(MorphGrid([]), eng.mus_anchor_host[a * 5 + b][np.newaxis])
        flat.allocate_ps_anchor(n)
        flat.ps_anchor[0].copy_(eng.ps_anchor[a * 5 + b])
        assert got[0] == flat.evaluate(np.zeros((1, 0)), mult[:1])[0]


@pytest.mark.parametrize("bb", [None, 0])
def test_config3_binned_full_size(bb):
    """3-D 200x200x20 bins, 4 sources, 3 shape parameters x 3 anchors (G = 27), Beeston-Barlow on source 0."""
    from blueice_b200.engine import BinnedEngine, MorphGrid
    axes, edges, mus, pmf, n_model, observed = wl.c3_arrays((200, 200, 20), 4, 3, (-1., 0., 1.), seed=3)
    grid = MorphGrid(axes)
    eng = BinnedEngine(grid, mus.reshape(27, 4), pmf, n_model if bb is not None else None, bb)
    eng.set_observed(observed)
    zs, mult = wl.scan_points(24, 3, 4, seed=4, z_range=(-1., 1.), mult_range=(0.8, 1.2))
    zs[0] = 0.0                                         # the base model (an anchor)
    zs[1] = [1.0, -1.0, 0.0]
    got, status, flags = eng.evaluate(zs, mult, return_status=True)
    assert np.all(status == 0) and np.all(flags == 0) and np.all(np.isfinite(got))
    for i in (0, 5, 23):                                # batch rows == single-point calls, bitwise
        assert eng.evaluate(zs[i:i + 1], mult[i:i + 1])[0] == got[i]
    orc = BinnedOracle(axes, mus, pmf, n_model if bb is not None else None, bb).set_observed(observed)
    n_bins = observed.size
    for i in (0, 1, 5, 9, 17, 23):
        ref = orc(zs[i], mult[i])
        assert abs(got[i] - ref) <= 1e-9 * max(observed.sum(), n_bins)
        assert abs(got[i] - ref) <= 1e-12 * (abs(ref) + n_bins), (got[i], ref)
    # device binning of 2e5 events into the same 8e5 bins reproduces np.histogramdd exactly
    cols = wl.events_from_counts(edges, observed, seed=1)
    assert np.array_equal(eng.histogram_events(edges, cols), observed)


# ------------------------------------------------------------------------------------------------
# configs 4 and 5 on the template-space engine (no dense anchor tensor), BASELINE shapes
# ------------------------------------------------------------------------------------------------
def test_config4_toy_mc_shape(monkeypatch):
    """3 sources, 3 shape parameters x 5 anchors (125 anchors), 100x100 templates, 1e5 toys x ~1000 events generated
    on the device (1e8 events; the full config is 10 such sweeps or 8 GPUs x 1.25e5 toys), one point per toy."""
    from blueice_b200 import toys as btoys
    from blueice_b200.engine import MorphGrid, TemplateUnbinnedEngine
    from oracle.pipeline import toy_loglikelihoods
    axes, edges, templates, mus = wl.c2_arrays(3, 3, wl.ANCHORS_5, (100, 100))
    grid = MorphGrid(axes)
    rows = templates.reshape((125 * 3, 100, 100))
    eng = TemplateUnbinnedEngine(grid, mus.reshape(125, 3), rows, edges)
    centre = (2, 2, 2)
    vol = np.outer(np.diff(edges[0]), np.diff(edges[1]))
    cdf = np.vstack([np.cumsum((templates[centre + (s,)] * vol).ravel()) for s in range(3)])
    cdf /= cdf[:, -1:]
    scale = 1000.0 / mus[centre].sum()
    T = 100000
    coords, source, offsets, counts = btoys.generate(edges, cdf, mus[centre] * scale, T, seed=7)
    assert abs(counts.sum(axis=1).mean() - 1000.0) < 1.0 and offsets[-1] == counts.sum()
    eng.set_datasets(coords, offsets)
    zs, mult = wl.scan_points(T, 3, 3, seed=4)
    sc = np.full(T, scale)
    got, status = eng.evaluate_toys(zs, mult, scale=sc, return_status=True)
    assert np.all(status == 0) and np.all(np.isfinite(got))
    # the device walks the toys in hypercube-cell order (k_ts_order_groups); in toy order (BI_TS_ORDER=0): same bits
    monkeypatch.setenv("BI_TS_ORDER", "0")
    assert np.array_equal(eng.evaluate_toys(zs, mult, scale=sc), got)
    monkeypatch.delenv("BI_TS_ORDER")
    # toy t through the single-dataset path (grouped schedule): same bits
    for t in (0, 1234, T - 1):
        assert eng.evaluate(zs[t:t + 1], mult[t:t + 1], scale=sc[:1], dataset=t)[0] == got[t]
    # the reference's own loop (oracle) on 64 toys
    pick = sorted(set([3, 50000, 99998] + list(np.random.default_rng(8).choice(T, size=61, replace=False))))
    sub_off = np.concatenate([[0], np.cumsum([offsets[t + 1] - offsets[t] for t in pick])])
    c_host = [np.concatenate([coords[k, offsets[t]:offsets[t + 1]].cpu().numpy() for t in pick]) for k in range(2)]
    want = toy_loglikelihoods(axes, mus * scale, templates, edges, c_host, sub_off, zs[pick], mult[pick])
    assert np.all(np.abs(got[pick] - want) <= 1e-9 * 1000)
    # toys generated in two pieces (two ranks) are the same toys: same likelihoods, bit for bit
    ca, _, oa, _ = btoys.generate(edges, cdf, mus[centre] * scale, 60000, seed=7)
    cb, _, ob, _ = btoys.generate(edges, cdf, mus[centre] * scale, 40000, seed=7, first_toy=60000)
    import torch
    eng.set_datasets(torch.cat([ca, cb], dim=1), np.concatenate([oa, oa[-1] + ob[1:]]))
    assert np.array_equal(eng.evaluate_toys(zs, mult, scale=sc), got)
    # on average a toy prefers the truth (base model) over a far-away point
    truth = eng.evaluate_toys(np.zeros((T, 3)), np.ones((T, 3)), scale=sc)
    far = eng.evaluate_toys(np.full((T, 3), 1.5), np.ones((T, 3)), scale=sc)
    assert truth.mean() > far.mean()


def test_config5_large_dataset_shape():
    """6 sources, 4 shape parameters x 5 anchors (625 anchors): 1.25e7 events (one GPU's share of the 1e8-event
    config) on the mixture engine -- batch-shape independence, event shards adding up, agreement with the exact
    template kernel and, on a sub-sample, with the anchor-tensor engine and the oracle."""
    import torch
    from blueice_b200 import toys as btoys
    from blueice_b200.engine import MorphGrid, TemplateUnbinnedEngine
    from oracle.pipeline import UnbinnedOracle
    axes, edges, templates, mus = wl.c2_arrays(6, 4, wl.ANCHORS_5, (100, 100))
    grid = MorphGrid(axes)
    rows = templates.reshape((625 * 6, 100, 100))
    centre = (2, 2, 2, 2)
    vol = np.outer(np.diff(edges[0]), np.diff(edges[1]))
    cdf = np.vstack([np.cumsum((templates[centre + (s,)] * vol).ravel()) for s in range(6)])
    cdf /= cdf[:, -1:]
    N_target = 12500000
    scale = N_target / mus[centre].sum()
    coords, _, offsets, _ = btoys.generate(edges, cdf, mus[centre] * scale, 1, seed=5)
    N = int(offsets[-1])
    mix = TemplateUnbinnedEngine(grid, mus.reshape(625, 6), rows, edges, mode='mixture')
    mix.set_datasets(coords)
    rng = np.random.default_rng(51)
    x0m, x0z = rng.uniform(0.8, 1.2, size=6), rng.uniform(-1.9, 1.9, size=4)
    zs = np.repeat(x0z[None], 11, 0)
    mult = np.repeat(x0m[None], 11, 0)
    for j in range(10):
        (mult if j < 6 else zs)[j + 1, j if j < 6 else j - 6] += 1.4901161193847656e-08
    sc = np.full(11, scale)
    got = mix.evaluate(zs, mult, scale=sc)
    assert np.all(np.isfinite(got))
    for p in (0, 4, 10):                                            # a finite-difference batch == single evaluations
        assert mix.evaluate(zs[p:p + 1], mult[p:p + 1], scale=sc[:1])[0] == got[p]
    # events split into 8 superblock-aligned shards (8 GPUs): the shards' log sums add up to the whole
    logsum, musum, _ = mix.evaluate(zs, mult, scale=sc, return_parts=True)
    from blueice_b200.distributed import shard_bounds
    bounds = shard_bounds(N, 8, align=512)
    offs = np.array([b[0] for b in bounds] + [N])
    mix.set_datasets(coords, offs)
    total = np.zeros(11)
    for r in range(8):
        total = total + mix.evaluate(zs, mult, scale=sc, return_parts=True, dataset=r)[0]
    assert np.all(np.abs(total - logsum) <= 1e-9 * N)
    assert np.all(np.abs(total - logsum) <= 1e-12 * (np.abs(logsum) + N))
    # exact template kernel on the same events
    exact = TemplateUnbinnedEngine(grid, mus.reshape(625, 6), rows, edges)
    exact.set_datasets(coords)
    ref = exact.evaluate(zs[:2], mult[:2], scale=sc[:2])
    assert np.all(np.abs(got[:2] - ref) <= 1e-9 * N) and np.all(np.abs(got[:2] - ref) <= 1e-12 * (np.abs(ref) + N))
    # oracle (the reference's dense path: a 3 GB anchor tensor on the host) on the first 1e5 events
    n_sub = 100000
    sub = coords[:, :n_sub].cpu().numpy()
    mix.set_datasets(sub)
    small = mix.evaluate(zs[:2], mult[:2], scale=sc[:2])
    want = UnbinnedOracle(axes, mus * scale).set_data_from_templates(templates, edges, list(sub)).batch(zs[:2], mult[:2])
    assert np.all(np.abs(small - want) <= 1e-9 * n_sub)
    assert np.all(np.abs(small - want) <= 1e-12 * (np.abs(want) + n_sub))
    del exact, mix
    torch.cuda.empty_cache()


def test_config5_full_size_mixture_against_the_exact_kernel():
    """The whole 1e8-event dataset of config 5 on ONE GPU: the mixture engine (K5b: templates morphed per point, events
    streamed) against the exact template kernel (K5, bit-identical to the anchor-tensor engine) on the same events, for a
    single point and inside a finite-difference batch; the difference between two nearby points (what a minimiser's
    gradient is made of) agrees too."""
    import torch
    from blueice_b200 import toys as btoys
    from blueice_b200.engine import MorphGrid, TemplateUnbinnedEngine
    axes, edges, templates, mus = wl.c2_arrays(6, 4, wl.ANCHORS_5, (100, 100))
    grid = MorphGrid(axes)
    rows = templates.reshape((625 * 6, 100, 100))
    centre = (2, 2, 2, 2)
    vol = np.outer(np.diff(edges[0]), np.diff(edges[1]))
    cdf = np.vstack([np.cumsum((templates[centre + (s,)] * vol).ravel()) for s in range(6)])
    cdf /= cdf[:, -1:]
    scale = 1.0e8 / mus[centre].sum()
    coords, _, offsets, _ = btoys.generate(edges, cdf, mus[centre] * scale, 1, seed=5)
    N = int(offsets[-1])
    assert 0.999e8 < N < 1.001e8
    rng = np.random.default_rng(52)
    zs = np.repeat(rng.uniform(-1.9, 1.9, size=4)[None], 11, 0)
    mult = np.repeat(rng.uniform(0.8, 1.2, size=6)[None], 11, 0)
    for j in range(10):
        (mult if j < 6 else zs)[j + 1, j if j < 6 else j - 6] += 1e-4
    sc = np.full(11, scale)
    mix = TemplateUnbinnedEngine(grid, mus.reshape(625, 6), rows, edges, mode='mixture')
    mix.set_datasets(coords)
    got = mix.evaluate(zs, mult, scale=sc)
    assert mix.evaluate(zs[:1], mult[:1], scale=sc[:1])[0] == got[0]
    del mix
    torch.cuda.empty_cache()
    exact = TemplateUnbinnedEngine(grid, mus.reshape(625, 6), rows, edges)
    exact.set_datasets(coords)
    ref = exact.evaluate(zs[[0, 3, 8]], mult[[0, 3, 8]], scale=sc[:3])
    diff = np.abs(got[[0, 3, 8]] - ref)
    assert np.all(diff <= 1e-9 * N), diff
    assert np.all(diff <= 1e-12 * (np.abs(ref) + N)), diff
    assert abs((got[3] - got[0]) - (ref[1] - ref[0])) <= 1e-9 * N
    del exact
    torch.cuda.empty_cache()
