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
    for mode in ('stream', 'grouped', None):
        eng.force_kernel = mode
        res[mode] = ll.batch(table, names)
    eng.force_kernel = None
    # kernel choice, batch composition and scalar calls: identical bits
    assert np.array_equal(res['stream'], res['grouped'])
    assert eng.plan(zs).kernel == 'mma'                  # the default path is the DMMA kernel
    dm = np.abs(res[None] - res['stream'])
    assert np.all(dm <= 2e-13 * (np.abs(res['stream']) + n)), dm.max()
    for i in (0, 1, 1000, 4095):
        assert ll(**dict(zip(names, [float(v) for v in table[i]]))) == res[None][i]
    perm = np.random.default_rng(0).permutation(4096)[:777]
    assert np.array_equal(ll.batch(table[perm], names), res[None][perm])
    # direct comparison with the oracle on a handful of points
    axes, edges, templates, mus = wl.c2_arrays(2, 2, wl.ANCHORS_5, (100, 100))
    orc = UnbinnedOracle(axes, mus).set_data_from_templates(templates, edges, [d['cs1'], d['cs2']])
    idx = [0, 7, 100, 2048, 4095]
    ref = orc.batch(zs[idx], mult[idx])
    diff = np.abs(res[None][idx] - ref)
    assert np.all(diff <= 1e-9 * n), diff
    assert np.all(diff <= 2e-13 * (np.abs(ref) + n)), diff
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
    for i in (0, 1, 9):
        ref = orc(zs[i], mult[i])
        assert abs(got[i] - ref) <= 1e-9 * max(observed.sum(), n_bins)
        assert abs(got[i] - ref) <= 1e-12 * (abs(ref) + n_bins), (got[i], ref)
    # device binning of 2e5 events into the same 8e5 bins reproduces np.histogramdd exactly
    cols = wl.events_from_counts(edges, observed, seed=1)
    assert np.array_equal(eng.histogram_events(edges, cols), observed)
