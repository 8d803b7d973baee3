"""Template-space engine (K5, blueice_b200/csrc/bi_template.cu): many datasets with one parameter point each
(toy Monte Carlos, BASELINE config 4) and single datasets without a dense anchor tensor (config 5).

Contract under test: a (dataset, point) pair evaluates BIT-IDENTICALLY to the anchor-tensor engine (K3 + K2) on
that dataset alone, and within |dlogL| <= 1e-9 * N of the oracle (the reference's own loop over the toys)."""
import numpy as np
import pytest

import bench_workloads as wl

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _model(n_sources, n_shape, anchors, bins):
    axes, edges, templates, mus = wl.c2_arrays(n_sources, n_shape, anchors, bins)
    return axes, edges, templates, mus


def _engines(axes, edges, templates, mus, method='linear', **kw):
    from blueice_b200.engine import MorphGrid, TemplateUnbinnedEngine, UnbinnedEngine
    grid = MorphGrid(axes)
    S = mus.shape[-1]
    n_bins = templates.shape[len(axes) + 1:]
    rows = templates.reshape((grid.n_anchors * S,) + tuple(n_bins))
    te = TemplateUnbinnedEngine(grid, mus.reshape(grid.n_anchors, S), rows, edges, method, **kw)
    ae = UnbinnedEngine(grid, mus.reshape(grid.n_anchors, S))
    return grid, te, ae, rows


def _anchor_eval(ae, rows, edges, coords, zs, mult, method='linear'):
    """The K3 + K2 path on one dataset."""
    import torch
    from blueice_b200 import _cabi
    n = coords.shape[1]
    ae.allocate_ps_anchor(n)
    if n:
        cd = torch.from_numpy(np.ascontiguousarray(coords)).to(ae.device)
        S = ae.n_sources
        ae.lookup_rows([(r // S, r % S) for r in range(len(rows))], rows, edges, cd,
                       _cabi.LOOKUP_LINEAR if method == 'linear' else _cabi.LOOKUP_PIECEWISE)
    return ae.evaluate(zs, mult)


@pytest.mark.parametrize("method", ["linear", "piecewise"])
@pytest.mark.parametrize("n_events", [0, 1, 31, 32, 33, 511, 512, 513, 1500])
def test_single_dataset_bit_identical_to_anchor_engine(method, n_events):
    axes, edges, templates, mus = _model(2, 2, (-1., 0., 1.), (40, 30))
    grid, te, ae, rows = _engines(axes, edges, templates, mus, method)
    rng = np.random.default_rng(n_events + 7)
    x, y = wl.c2_events(templates, mus, edges, max(n_events, 1), seed=5)
    coords = np.vstack([x, y])[:, :n_events]
    zs, mult = wl.scan_points(70, 2, 2, seed=3, z_range=(-1., 1.))
    zs[5] = [1.5, 0.]                       # out of range -> -inf
    zs[6] = [-1., 1.]                       # on the grid boundary
    zs[7] = [0., 0.]                        # on an anchor
    mult[8, 0] = -1.0                       # unphysical -> -inf
    te.set_datasets(coords)
    got = te.evaluate(zs, mult)
    ref = _anchor_eval(ae, rows, edges, coords, zs, mult, method)
    assert np.array_equal(got, ref)
    assert np.isneginf(got[5]) and np.isneginf(got[8]) and np.isfinite(got[6])
    one = te.evaluate(zs[:1], mult[:1])     # P = 1 uses the ungrouped kernel: same bits
    assert one[0] == got[0]


def test_packed_layout_holds_the_lookup_neighbours():
    """K5's template layout for linear lookups: every bin with its neighbours along the last two dimensions."""
    axes, edges, templates, mus = _model(3, 3, (-1., 0., 1.), (25, 20))
    _, te, _, rows = _engines(axes, edges, templates, mus)
    flat = rows.reshape(len(rows), -1)
    packed = te.templates.cpu().numpy()
    assert packed.shape == (len(rows), 500, 4) and te.row_stride == 2000 and te.bin_stride == 4
    assert np.array_equal(packed[:, :, 0], flat) and np.array_equal(packed[:, :-1, 1], flat[:, 1:])
    assert np.array_equal(packed[:, :-20, 2], flat[:, 20:]) and np.array_equal(packed[:, :-21, 3], flat[:, 21:])
    _, tp, _, _ = _engines(axes, edges, templates, mus, 'piecewise')
    assert tp.templates.shape == (len(rows), 500) and tp.bin_stride == 1


def test_zero_density_events_take_the_reference_outlier_path():
    """Templates with empty bins: events there have p = 0 -> outlier_likelihood (likelihood.py:686-689)."""
    axes, edges, templates, mus = _model(2, 1, (-1., 0., 1.), (20, 20))
    templates = templates.copy()
    templates[..., :6, :] = 0.0             # a dead region in every template
    grid, te, ae, rows = _engines(axes, edges, templates, mus)
    rng = np.random.default_rng(2)
    coords = np.vstack([rng.uniform(0, 100, 2000), rng.uniform(0, 4, 2000)])
    zs, mult = wl.scan_points(20, 1, 2, seed=8, z_range=(-1., 1.))
    te.set_datasets(coords)
    got = te.evaluate(zs, mult)
    ref = _anchor_eval(ae, rows, edges, coords, zs, mult)
    assert np.array_equal(got, ref)
    from oracle.pipeline import UnbinnedOracle
    orc = UnbinnedOracle(axes, mus).set_data_from_templates(templates, edges, list(coords))
    want = orc.batch(zs, mult)
    assert np.all(np.abs(got - want) <= TOL * 2000)


@pytest.mark.parametrize("n_space", [1, 3])
def test_other_analysis_dimensionalities(n_space):
    """1-D and 3-D templates use scipy's generic corner loop (not the 2-D fast path)."""
    from blueice_b200.engine import MorphGrid, TemplateUnbinnedEngine, UnbinnedEngine
    from oracle.pipeline import UnbinnedOracle
    rng = np.random.default_rng(n_space)
    axes = [np.array([-1., 0., 2.])]
    n_bins = (17, 9, 5)[:n_space]
    edges = [np.linspace(0., 1. + d, n + 1) for d, n in enumerate(n_bins)]
    templates = rng.uniform(0.2, 1.5, size=(3, 2) + n_bins)
    mus = rng.uniform(50., 100., size=(3, 2))
    n = 777
    coords = np.vstack([rng.uniform(-0.1, 1.1 + d, n) for d in range(n_space)])       # some outside: clipped
    grid = MorphGrid(axes)
    te = TemplateUnbinnedEngine(grid, mus, templates.reshape((6,) + n_bins), edges)
    te.set_datasets(coords)
    zs = rng.uniform(-1., 2., size=(33, 1))
    mult = rng.uniform(0.5, 1.5, size=(33, 2))
    got = te.evaluate(zs, mult)
    ae = UnbinnedEngine(grid, mus)
    ref = _anchor_eval(ae, templates.reshape((6,) + n_bins), edges, coords, zs, mult)
    assert np.array_equal(got, ref)
    want = UnbinnedOracle(axes, mus).set_data_from_templates(templates, edges, list(coords)).batch(zs, mult)
    assert np.all(np.abs(got - want) <= TOL * n)


def test_toys_equal_per_toy_evaluation_and_oracle():
    """Ragged toys (incl. empty ones), one point each == set_data(toy) + ll(point) bit for bit; oracle within 1e-9 N."""
    from oracle.pipeline import toy_loglikelihoods
    axes, edges, templates, mus = _model(3, 3, (-1., 0., 1.), (30, 24))
    grid, te, ae, rows = _engines(axes, edges, templates, mus)
    rng = np.random.default_rng(11)
    T = 60
    sizes = rng.poisson(700, size=T)
    sizes[3] = 0
    sizes[10] = 1
    sizes[11] = 512
    sizes[12] = 1025
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    x, y = wl.c2_events(templates, mus, edges, int(offsets[-1]) + 10, seed=21)
    coords = np.vstack([x, y])[:, :offsets[-1]]
    zs, mult = wl.scan_points(T, 3, 3, seed=6, z_range=(-1., 1.))
    zs[20] = [0., 2., 0.]                   # out of range
    te.set_datasets(coords, offsets)
    got, status = te.evaluate_toys(zs, mult, return_status=True)
    for t in range(T):
        ref = _anchor_eval(ae, rows, edges, coords[:, offsets[t]:offsets[t + 1]], zs[t:t + 1], mult[t:t + 1])
        assert got[t] == ref[0], (t, got[t], ref[0])
    assert np.isneginf(got[20]) and status[20] != 0
    want = toy_loglikelihoods(axes, mus, templates, edges, list(coords), offsets, zs, mult)
    fin = np.isfinite(want)
    assert np.array_equal(fin, np.isfinite(got))
    assert np.all(np.abs(got[fin] - want[fin]) <= TOL * np.maximum(sizes[fin], 1))
    # every toy on dataset-level evaluate(): same numbers through the grouped schedule
    again = np.array([te.evaluate(zs[t:t + 1], mult[t:t + 1], dataset=t)[0] for t in (0, 5, 12)])
    assert np.array_equal(again, got[[0, 5, 12]])


def test_api_template_engine_and_toys():
    """likelihood_config['unbinned_engine'] = 'template' and set_toy_data / batch_toys through the class API."""
    ll_a, d, names = wl.c2_api(n_sources=2, n_shape=2, anchors=(-1., 0., 1.), bins=(40, 30), n_events=3000, seed=3)
    ll_t, d2, _ = wl.c2_api(n_sources=2, n_shape=2, anchors=(-1., 0., 1.), bins=(40, 30), n_events=3000, seed=3,
                            likelihood_config={'unbinned_engine': 'template'})
    from blueice_b200.engine import TemplateUnbinnedEngine
    assert isinstance(ll_t._engine, TemplateUnbinnedEngine)
    zs, mult = wl.scan_points(100, 2, 2, seed=12, z_range=(-1.2, 1.2))
    params = np.column_stack([mult, zs])
    assert np.array_equal(ll_t.batch(params, names), ll_a.batch(params, names))
    kw = dict(zip(names, [float(v) for v in params[1]]))
    assert ll_t(**kw) == ll_a(**kw)
    r_t, mus_t, ps_t = ll_t(full_output=True, **kw)
    r_a, mus_a, ps_a = ll_a(full_output=True, **kw)
    assert r_t == r_a and np.array_equal(mus_t, mus_a) and np.array_equal(ps_t, ps_a)

    # toys: slices of d as separate datasets, one point each
    cuts = [0, 700, 700, 1900, len(d)]
    toys = [d[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    ll_a.set_toy_data(toys)
    got = ll_a.batch_toys(params[:4], names)
    for t, toy in enumerate(toys):
        ll_a.set_data(toy)
        assert got[t] == ll_a(**dict(zip(names, [float(v) for v in params[t]])))
    # the same toys as one concatenated dataset + offsets
    ll_a.set_toy_data(d, offsets=np.asarray(cuts))
    assert np.array_equal(ll_a.batch_toys(params[:4], names), got)
    with pytest.raises(ValueError):
        ll_a.batch_toys(params[:3], names)


# ------------------------------------------------------------------------------------------------
# mixture form (K5b): morph the templates, look the events up in the mixture template
# ------------------------------------------------------------------------------------------------
def assert_close(got, ref, n_events):
    assert np.array_equal(np.isneginf(got), np.isneginf(ref))
    fin = np.isfinite(ref)
    err = np.abs(got[fin] - ref[fin])
    assert np.all(err <= TOL * max(n_events, 1)), err.max()
    assert np.all(err <= 1e-12 * (np.abs(ref[fin]) + n_events)), err.max()       # what the arithmetic really delivers


@pytest.mark.parametrize("method", ["linear", "piecewise"])
@pytest.mark.parametrize("n_events", [0, 1, 33, 512, 1500, 20000])
def test_mixture_form_matches_exact_form_and_oracle(method, n_events):
    from oracle.pipeline import UnbinnedOracle
    axes, edges, templates, mus = _model(2, 2, (-1., 0., 1.), (40, 30))
    _, te, _, _ = _engines(axes, edges, templates, mus, method)
    _, tm, _, _ = _engines(axes, edges, templates, mus, method, mode='mixture')
    x, y = wl.c2_events(templates, mus, edges, max(n_events, 1), seed=5)
    coords = np.vstack([x, y])[:, :n_events]
    zs, mult = wl.scan_points(37, 2, 2, seed=3, z_range=(-1., 1.))
    zs[5] = [1.5, 0.]
    zs[7] = [0., 0.]
    mult[8, 0] = -1.0
    te.set_datasets(coords)
    tm.set_datasets(coords)
    exact = te.evaluate(zs, mult)
    got, status = tm.evaluate(zs, mult, return_status=True)
    assert_close(got, exact, n_events)
    assert status[5] != 0 and status[8] != 0 and np.isneginf(got[5])
    assert tm.evaluate(zs[:1], mult[:1])[0] == got[0]                 # batch-shape independent, bitwise
    assert np.array_equal(tm.evaluate(zs[::-1], mult[::-1])[::-1], got)
    if n_events <= 1500:
        want = UnbinnedOracle(axes, mus).set_data_from_templates(templates, edges, list(coords), method).batch(zs, mult)
        assert_close(got, want, n_events)
    # event-sharded terms: logsum + musum reproduce logl
    logsum, musum, st = tm.evaluate(zs, mult, return_parts=True)
    ok = st == 0
    assert np.array_equal(-musum[ok] + logsum[ok], got[ok])


@pytest.mark.parametrize("n_space", [1, 2, 3, 4])
@pytest.mark.parametrize("n_events", [40, 1000, 30000])
def test_grouped_mixture_kernel_equals_single_point_kernel_bitwise(n_space, n_events):
    """Groups of up to 8 points run on the FP64 tensor pipe (k_mixture_partials_mma: one 8 x 8 x 4 contraction per octet
    of events that share a bin, eight when they straddle a bin edge); a single point runs on the gather kernel.  Same
    bits, for few events per bin (mostly straddling octets), many events per bin (mostly shared bins), empty bins
    (p = 0 -> outlier path) and several ragged datasets."""
    from blueice_b200.engine import MorphGrid, TemplateUnbinnedEngine
    from oracle.pipeline import UnbinnedOracle
    rng = np.random.default_rng(10 * n_space + 1)
    axes = [np.array([-1., 0., 2.])]
    n_bins = (23, 9, 5, 3)[:n_space]
    edges = [np.linspace(0., 1. + d, n + 1) for d, n in enumerate(n_bins)]
    templates = rng.uniform(0.2, 1.5, size=(3, 2) + n_bins)
    templates[:, :, :3] = 0.0                                           # a dead region in every template
    mus = rng.uniform(50., 100., size=(3, 2))
    coords = np.vstack([rng.uniform(-0.1, 1.1 + d, n_events) for d in range(n_space)])
    offsets = np.array([0, n_events // 3 + 1, n_events // 3 + 1, n_events])
    grid = MorphGrid(axes)
    tm = TemplateUnbinnedEngine(grid, mus, templates.reshape((6,) + n_bins), edges, mode='mixture')
    tm.mix_wide_min_superblocks = 0                                     # groups of 16 whatever the dataset size
    tm.set_datasets(coords, offsets)
    zs = rng.uniform(-1., 2., size=(29, 1))
    zs[3] = 2.5                                                         # out of range -> -inf inside a group
    mult = rng.uniform(0.5, 1.5, size=(29, 2))
    for t in range(3):
        got = tm.evaluate(zs, mult, dataset=t)
        single = np.array([tm.evaluate(zs[i:i + 1], mult[i:i + 1], dataset=t)[0] for i in range(len(zs))])
        assert np.array_equal(got, single)                              # groups of 16: two 8-point m-tiles per warp
        assert np.array_equal(tm.evaluate(zs[:7], mult[:7], dataset=t), single[:7])      # one m-tile
        assert np.isneginf(got[3]) and np.all(np.isfinite(np.delete(got, 3)))
        if n_events <= 1000:
            sl = slice(offsets[t], offsets[t + 1])
            want = UnbinnedOracle(axes, mus).set_data_from_templates(templates, edges, list(coords[:, sl])).batch(zs, mult)
            assert_close(got, want, sl.stop - sl.start)


def test_mixture_form_outlier_semantics_and_datasets():
    """Dead template regions (p = 0 -> outlier_likelihood) and several datasets on one engine."""
    axes, edges, templates, mus = _model(2, 1, (-1., 0., 1.), (20, 20))
    templates = templates.copy()
    templates[..., :6, :] = 0.0
    _, te, _, _ = _engines(axes, edges, templates, mus)
    _, tm, _, _ = _engines(axes, edges, templates, mus, mode='mixture')
    rng = np.random.default_rng(2)
    coords = np.vstack([rng.uniform(0, 100, 3000), rng.uniform(0, 4, 3000)])
    offsets = np.array([0, 1000, 1000, 3000])
    zs, mult = wl.scan_points(20, 1, 2, seed=8, z_range=(-1., 1.))
    te.set_datasets(coords, offsets)
    tm.set_datasets(coords, offsets)
    for t, n in enumerate(np.diff(offsets)):
        assert_close(tm.evaluate(zs, mult, dataset=t), te.evaluate(zs, mult, dataset=t), n)
    with pytest.raises(NotImplementedError):
        tm.evaluate_toys(zs[:3], mult[:3])
    bad = templates.copy()
    bad[0, 0, 10, 10] = np.nan
    with pytest.raises(ValueError):
        _engines(axes, edges, bad, mus, mode='mixture')


def test_api_mixture_engine():
    ll_a, d, names = wl.c2_api(n_sources=2, n_shape=2, anchors=(-1., 0., 1.), bins=(40, 30), n_events=3000, seed=3)
    ll_m, _, _ = wl.c2_api(n_sources=2, n_shape=2, anchors=(-1., 0., 1.), bins=(40, 30), n_events=3000, seed=3,
                           likelihood_config={'unbinned_engine': 'mixture'})
    assert ll_m._engine.mode == 'mixture'
    zs, mult = wl.scan_points(100, 2, 2, seed=12, z_range=(-1.2, 1.2))
    params = np.column_stack([mult, zs])
    assert_close(ll_m.batch(params, names), ll_a.batch(params, names), len(d))
    kw = dict(zip(names, [float(v) for v in params[1]]))
    r_m, mus_m, ps_m = ll_m(full_output=True, **kw)
    r_a, mus_a, ps_a = ll_a(full_output=True, **kw)
    assert abs(r_m - r_a) <= TOL * len(d) and np.array_equal(mus_m, mus_a) and np.array_equal(ps_m, ps_a)
    # 'auto' keeps the dense anchor tensor for a dataset this small
    ll_auto, _, _ = wl.c2_api(n_sources=2, n_shape=2, anchors=(-1., 0., 1.), bins=(40, 30), n_events=300, seed=3,
                              likelihood_config={'unbinned_engine': 'auto'})
    from blueice_b200.engine import UnbinnedEngine
    assert type(ll_auto._engine) is UnbinnedEngine


@pytest.mark.parametrize("mode", ["exact", "mixture"])
def test_one_call_entry_equals_the_staged_calls(mode):
    """bi_template_ll_batch (K1 -> morph -> K5 / K5b -> finalize in one C call) == the same stages called one by one
    (bi_point_setup, bi_template_mix, bi_template_partials / bi_mixture_partials, bi_template_finalize), bit for bit."""
    axes, edges, templates, mus = _model(3, 2, (-1., 0., 1.), (30, 24))
    _, te, _, _ = _engines(axes, edges, templates, mus, mode=mode)
    x, y = wl.c2_events(templates, mus, edges, 5000, seed=9)
    offsets = np.array([0, 1200, 1200, 3300, 5000])
    te.set_datasets(np.vstack([x, y])[:, :5000], offsets)
    zs, mult = wl.scan_points(23, 2, 3, seed=4, z_range=(-1.1, 1.1))
    datasets = np.arange(23) % 4
    sched, order = te.pair_schedule(datasets, zs)
    zs_d, mult_d, scale_d, eff_d, _ = te._upload_points(zs, mult, None, None)
    zs_d, mult_d = zs_d.clone(), mult_d.clone()
    o1, logl1, logsum1 = te.run_one_call(23, sched, zs_d, mult_d, None, None)
    a = (logl1.cpu().numpy().copy(), logsum1.cpu().numpy().copy(), o1["musum"].cpu().numpy().copy(),
         o1["status"].cpu().numpy().copy())
    o2 = te._setup_terms(23, zs_d, mult_d, None, None)
    logl2, logsum2 = te.run_schedule(sched, o2)
    b = (logl2.cpu().numpy(), logsum2.cpu().numpy(), o2["musum"].cpu().numpy(), o2["status"].cpu().numpy())
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    assert np.any(a[3] != 0) and np.any(np.isneginf(a[0]))           # out-of-range points included


@pytest.mark.parametrize("shape", [(3, 3, (30, 24)), (2, 1, (40,)), (5, 2, (25, 20)), (1, 4, (12, 10))])
def test_bin_major_toy_sweep_is_bit_identical_to_k5(shape, monkeypatch):
    """K5c (bi_template_bm.cu): the densities of a toy sweep formed bin-major == the toy-by-toy kernel, bit for bit --
    ragged and empty toys, out-of-range and unphysical points, dead template regions (rare path), 1-D and 2-D templates,
    1..4 shape parameters."""
    n_sources, n_shape, bins = shape
    anchors = (-1., 0., 1.) if n_shape < 4 else (-1., 1.)
    if len(bins) == 2:
        axes, edges, templates, mus = _model(n_sources, n_shape, anchors, bins)
    else:
        rng0 = np.random.default_rng(3)
        axes = [np.asarray(anchors)] * n_shape
        edges = [np.linspace(0., 2., bins[0] + 1)]
        g = len(anchors) ** n_shape
        templates = rng0.uniform(0.2, 1.5, size=(len(anchors),) * n_shape + (n_sources,) + bins)
        mus = rng0.uniform(50., 100., size=(len(anchors),) * n_shape + (n_sources,))
    templates = templates.copy()
    templates[..., :3] = 0.0                                    # a dead strip: p = 0 -> outlier path in some toys
    rng = np.random.default_rng(17)
    T = 300
    sizes = rng.poisson(40, size=T)
    sizes[[3, 77]] = 0
    sizes[10] = 1
    sizes[12] = 1100
    sizes[13] = 5000                                            # one bin-major task cannot hold all events of a bin
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    n = int(offsets[-1])
    coords = np.vstack([rng.uniform(e[0] - 0.05, e[-1] + 0.05, n) for e in edges])            # some outside: clipped
    coords[:, offsets[13]:offsets[13] + 3000] = coords[:, offsets[13]:offsets[13] + 1]       # 3000 events in ONE bin
    zs, mult = wl.scan_points(T, n_shape, n_sources, seed=6, z_range=(-1., 1.))
    zs[20, 0] = 2.0                                             # out of range
    mult[21, 0] = -1.0                                          # unphysical
    zs[22] = -1.0                                               # grid corner
    zs[23] = 0.0 if n_shape < 4 else 1.0                        # on an anchor

    monkeypatch.setenv("BI_TS_BM", "0")
    _, te0, _, _ = _engines(axes, edges, templates, mus)
    te0.set_datasets(coords, offsets)
    want, st0 = te0.evaluate_toys(zs, mult, return_status=True)
    assert te0.toy_schedule()["bm"] is None
    monkeypatch.setenv("BI_TS_BM", "1")
    _, te1, _, _ = _engines(axes, edges, templates, mus)
    te1.set_datasets(coords, offsets)
    assert te1.toy_schedule()["bm"] is not None
    for _ in range(4):                                          # eager calls, then the CUDA-graph replay
        got, st1 = te1.evaluate_toys(zs, mult, return_status=True)
        assert np.array_equal(got, want) and np.array_equal(st0, st1)
    assert np.isneginf(got[20]) and np.isneginf(got[21]) and np.isfinite(got[22]) and np.isfinite(got[23])
    ls0, mu0, _ = te0.evaluate_toys(zs, mult, return_parts=True)
    ls1, mu1, _ = te1.evaluate_toys(zs, mult, return_parts=True)
    assert np.array_equal(ls0, ls1) and np.array_equal(mu0, mu1)
    # other points on the same toys (the records are rebuilt per call)
    zs2, mult2 = wl.scan_points(T, n_shape, n_sources, seed=9, z_range=(-1., 1.))
    assert np.array_equal(te1.evaluate_toys(zs2, mult2), te0.evaluate_toys(zs2, mult2))
