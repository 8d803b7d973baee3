"""On-device toy generation (blueice_b200/csrc/bi_toys.cu): the event stage is
bit-identical to oracle/toys.py (Philox pinned by Random123 vectors); the Poisson stage and the overall
distribution are checked statistically against the reference's rules (model.py:69-91, source.py:248-264)."""
import numpy as np
import pytest

import bench_workloads as wl

pytestmark = pytest.mark.gpu


def _tables(n_sources=3, bins=(30, 20)):
    axes, edges, templates, mus = wl.c2_arrays(n_sources, 1, (-1., 0., 1.), bins)
    vol = np.outer(np.diff(edges[0]), np.diff(edges[1]))
    pmf = templates[1] * vol
    cdf = np.vstack([np.cumsum(p.ravel()) / p.sum() for p in pmf])
    return edges, pmf, cdf


def test_events_bit_identical_to_oracle_and_shard_invariant():
    from blueice_b200 import toys
    from oracle import toys as otoys
    edges, pmf, cdf = _tables()
    mus = np.array([40., 7., 0.5])
    coords, source, offsets, counts = toys.generate(edges, cdf, mus, 300, seed=1234567890123, first_toy=5)
    want_c, want_s, want_o = otoys.toy_events(edges, cdf, counts, seed=1234567890123, first_toy=5)
    assert np.array_equal(offsets, want_o)
    assert np.array_equal(source.cpu().numpy(), want_s)
    assert np.array_equal(coords.cpu().numpy(), want_c)
    # the same toys generated in two pieces (as two ranks would)
    c_a, s_a, o_a, n_a = toys.generate(edges, cdf, mus, 120, seed=1234567890123, first_toy=5)
    c_b, s_b, o_b, n_b = toys.generate(edges, cdf, mus, 180, seed=1234567890123, first_toy=125)
    assert np.array_equal(np.vstack([n_a, n_b]), counts)
    assert np.array_equal(np.hstack([c_a.cpu().numpy(), c_b.cpu().numpy()]), want_c)
    # a different seed gives different toys
    c2 = toys.generate(edges, cdf, mus, 300, seed=1, first_toy=5)[3]
    assert not np.array_equal(c2, counts)


@pytest.mark.parametrize("mu", [0.0, 0.3, 4.0, 9.99, 10.0, 37.5, 1000.0, 250000.0])
def test_poisson_counts_distribution(mu):
    """Both NumPy-legacy algorithms (multiplication below 10, PTRS above): mean, variance and, for small means,
    the pmf itself."""
    from scipy import stats
    from blueice_b200 import toys
    edges, pmf, cdf = _tables(1)
    T = 200000
    import torch
    from blueice_b200 import _cabi
    import ctypes
    lib = _cabi.load()
    dev = torch.device("cuda:0")
    mus_d = torch.tensor([mu], dtype=torch.float64, device=dev)
    counts_d = torch.empty((T, 1), dtype=torch.int32, device=dev)
    _cabi.check(lib.bi_toy_counts(1, T, 0, _cabi.dev_ptr(mus_d), 0, 99, _cabi.dev_ptr(counts_d),
                                  ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "bi_toy_counts")
    k = counts_d.cpu().numpy().ravel().astype(np.int64)
    if mu == 0.0:
        assert np.all(k == 0)
        return
    assert np.all(k >= 0)
    # mean within 5 sigma; variance within 5 sigma of its sampling error
    assert abs(k.mean() - mu) <= 5 * np.sqrt(mu / T)
    var_err = np.sqrt((mu + 2 * mu * mu) / T)           # sd of the sample variance of a Poisson (approx.)
    assert abs(k.var() - mu) <= 5 * var_err
    if mu <= 40:
        top = int(mu + 8 * np.sqrt(mu) + 10)
        obs = np.bincount(np.minimum(k, top), minlength=top + 1).astype(float)
        exp = stats.poisson(mu).pmf(np.arange(top + 1)) * T
        exp[top] = stats.poisson(mu).sf(top - 1) * T
        keep = exp >= 5
        chi2 = np.sum((obs[keep] - exp[keep]) ** 2 / exp[keep]) + (obs[~keep].sum() - exp[~keep].sum()) ** 2 / max(exp[~keep].sum(), 1e-9)
        assert chi2 <= stats.chi2(keep.sum()).ppf(1 - 1e-6), chi2


def test_event_distribution_follows_the_source_pmfs():
    from scipy import stats
    from blueice_b200 import toys
    edges, pmf, cdf = _tables()
    mus = np.array([600., 300., 100.])
    coords, source, offsets, counts = toys.generate(edges, cdf, mus, 2000, seed=3)
    c = coords.cpu().numpy()
    s = source.cpu().numpy()
    assert np.all((c[0] >= edges[0][0]) & (c[0] <= edges[0][-1]) & (c[1] >= edges[1][0]) & (c[1] <= edges[1][-1]))
    for k in range(3):
        m = s == k
        h, _, _ = np.histogram2d(c[0][m], c[1][m], bins=edges)
        exp = pmf[k] / pmf[k].sum() * m.sum()
        keep = exp >= 5
        chi2 = np.sum((h[keep] - exp[keep]) ** 2 / exp[keep])
        assert chi2 <= stats.chi2(keep.sum() - 1).ppf(1 - 1e-6), (k, chi2, keep.sum())
        # uniform inside the bins
        frac = (c[0][m] - edges[0][np.searchsorted(edges[0], c[0][m], 'right') - 1]) / np.diff(edges[0])[0]
        assert abs(frac.mean() - 0.5) <= 5 * np.sqrt(1 / 12 / m.sum())
    # toys are grouped by source in source order
    t = 17
    seg = s[offsets[t]:offsets[t + 1]]
    assert np.all(np.diff(seg) >= 0) and np.array_equal(np.bincount(seg, minlength=3), counts[t])


def test_model_simulate_toys_feeds_batch_toys():
    """Model.simulate_toys -> set_toy_data -> batch_toys == per-toy set_data + call, bit for bit; toy means as in
    Model.simulate (rate multipliers, livetime)."""
    ll, d, names = wl.c2_api(n_sources=2, n_shape=2, anchors=(-1., 0., 1.), bins=(40, 30), n_events=500, seed=3)
    model = ll.base_model
    td = model.simulate_toys(40, rate_multipliers={'sig': 2.0}, livetime_days=0.01, seed=5)
    assert len(td) == 40 and td.n_events == td.offsets[-1]
    expect = np.array([model.expected_events(s) for s in model.sources]) * np.array([1.0, 2.0]) * 0.01
    assert abs(td.counts[:, 0].mean() - expect[0]) <= 5 * np.sqrt(expect[0] / 40)
    assert abs(td.counts[:, 1].mean() - expect[1]) <= 5 * np.sqrt(expect[1] / 40)
    ll.set_toy_data(td)
    zs, mult = wl.scan_points(40, 2, 2, seed=12, z_range=(-1., 1.))
    params = np.column_stack([mult, zs])
    got = ll.batch_toys(params, names)
    for t in (0, 7, 39):
        rec = td.to_records(t)
        assert len(rec) == td.counts[t].sum() and set(rec.dtype.names) == {'source', 'cs1', 'cs2'}
        ll.set_data(rec)
        assert got[t] == ll(**dict(zip(names, [float(v) for v in params[t]])))


def test_toy_index_evaluates_any_point_on_any_toy():
    ll, d, names = wl.c2_api(n_sources=2, n_shape=2, anchors=(-1., 0., 1.), bins=(40, 30), n_events=500, seed=3)
    td = ll.base_model.simulate_toys(30, livetime_days=0.01, seed=6)
    ll.set_toy_data(td)
    zs, mult = wl.scan_points(30, 2, 2, seed=12, z_range=(-1., 1.))
    params = np.column_stack([mult, zs])
    base = ll.batch_toys(params, names, livetime_days=0.01)
    # several points per toy, toys in arbitrary order; finite-difference neighbours share the group of their toy
    toy_index = np.array([5, 5, 5, 29, 0, 5, 17, 0])
    rows = params[[5, 6, 7, 29, 0, 5, 17, 3]].copy()
    rows[1, 0] += 1e-8
    got = ll.batch_toys(rows, names, livetime_days=0.01, toy_index=toy_index)
    assert got[0] == base[5] and got[5] == base[5] and got[3] == base[29] and got[4] == base[0] and got[6] == base[17]
    for q in (1, 2, 7):
        ll.set_data(td.to_records(int(toy_index[q])))
        assert got[q] == ll(livetime_days=0.01, **dict(zip(names, [float(v) for v in rows[q]])))
    with pytest.raises(ValueError):
        ll.batch_toys(rows, names, toy_index=toy_index[:3])
    with pytest.raises(ValueError):
        ll.batch_toys(rows[:1], names, toy_index=[30])


def test_bestfit_toys_matches_per_toy_scipy_fits():
    """All toys fitted in lock step on the device reach the optimum scipy finds toy by toy (the reference's way)."""
    from blueice_b200.inference import bestfit_scipy, bestfit_toys
    ll, d, names = wl.c2_api(n_sources=2, n_shape=2, anchors=(-2., -1., 0., 1., 2.), bins=(40, 30), n_events=500, seed=3)
    lt = 0.02
    td = ll.base_model.simulate_toys(24, livetime_days=lt, seed=8)
    ll.set_toy_data(td)
    fit, maxll, info = bestfit_toys(ll, livetime_days=lt)
    assert set(fit.keys()) == set(names) and maxll.shape == (24,)
    assert info['converged'].all(), info
    truth = ll.batch_toys(np.tile([1., 1., 0., 0.], (24, 1)), names, livetime_days=lt)
    assert np.all(maxll >= truth)                                    # a fit is at least as good as the truth
    for t in (0, 11, 23):
        ll.set_data(td.to_records(t))
        ref_fit, ref_ll = bestfit_scipy(ll, pass_bounds_to_minimizer=True,
                                        minimize_kwargs=dict(method='L-BFGS-B'), livetime_days=lt)
        assert abs(maxll[t] - ref_ll) <= 5e-2, (t, maxll[t], ref_ll)     # kinks at the anchors: neighbouring local optima
        got_ll = ll(livetime_days=lt, **{n: float(fit[n][t]) for n in names})
        assert got_ll == maxll[t]                                    # the reported maximum is the likelihood at the fit
    # a fixed parameter stays fixed (conditional fits of a profile-likelihood test statistic)
    cfit, cll, _ = bestfit_toys(ll, livetime_days=lt, sig_rate_multiplier=0.0)
    assert 'sig_rate_multiplier' not in cfit and np.all(cll <= maxll + 1e-6)


# ------------------------------------------------------------------------------------------------
# binned toys: every toy binned on the device, K4 with one observed histogram per point
# ------------------------------------------------------------------------------------------------
def _binned_model(bb):
    from blueice_b200.engine import BinnedEngine, MorphGrid
    axes, edges, mus, pmf, n_model, observed = wl.c3_arrays((12, 10, 4), 3, 2, (-1., 0., 1.), seed=3, total_events=3000.)
    grid = MorphGrid(axes)
    eng = BinnedEngine(grid, mus.reshape(9, 3), pmf, n_model if bb is not None else None, bb)
    return eng, edges, mus, pmf


@pytest.mark.parametrize("bb", [None, 0])
def test_binned_toys_equal_per_toy_evaluation(bb):
    eng, edges, mus, pmf = _binned_model(bb)
    rng = np.random.default_rng(5)
    T = 17
    sizes = rng.poisson(2500, size=T)
    sizes[4] = 0
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    n = int(offsets[-1])
    coords = np.vstack([rng.uniform(-6.5, 6.5, n), rng.uniform(-6, 6, n), rng.uniform(0, 1, n)])   # some outside: dropped
    coords[0, 7] = 6.0                                                  # on the right-most edge: last bin
    eng.set_observed_toys(edges, coords, offsets)
    # device binning of every toy == np.histogramdd of that toy
    for t in (0, 4, T - 1):
        ref, _ = np.histogramdd(coords[:, offsets[t]:offsets[t + 1]].T, bins=edges)
        assert np.array_equal(eng.toy_observed[t, :eng.n_bins].cpu().numpy().reshape(ref.shape), ref)
    zs, mult = wl.scan_points(T, 2, 3, seed=4, z_range=(-1., 1.), mult_range=(0.8, 1.2))
    zs[2] = [1.5, 0.]                                                   # out of range -> -inf
    got, status, flags = eng.evaluate_toys(zs, mult, return_status=True)
    assert np.isneginf(got[2]) and status[2] != 0
    for t in range(T):
        ref, _ = np.histogramdd(coords[:, offsets[t]:offsets[t + 1]].T, bins=edges)
        eng.set_observed(ref)
        one, st1, fl1 = eng.evaluate(zs[t:t + 1], mult[t:t + 1], return_status=True)
        assert (one[0] == got[t]) or (np.isneginf(one[0]) and np.isneginf(got[t])), (t, one[0], got[t])
        assert fl1[0] == flags[t]
    with pytest.raises(ValueError):
        eng.evaluate_toys(zs[:3], mult[:3])


def test_binned_toys_through_the_api():
    """BinnedLogLikelihood.set_toy_data(Model.simulate_toys(...)) + batch_toys == set_data(toy) + call per toy."""
    from blueice_b200 import BinnedLogLikelihood, HistogramPdfSource
    ll_u, d, names = wl.c2_api(n_sources=2, n_shape=2, anchors=(-1., 0., 1.), bins=(20, 15), n_events=400, seed=3)
    config = ll_u.pdf_base_config
    ll = BinnedLogLikelihood(config)
    for spec in config['sources']:
        ll.add_rate_parameter(spec['name'])
    for n in ('shift1', 'shift2'):
        ll.add_shape_parameter(n, (-1., 0., 1.))
    ll.prepare()
    td = ll.base_model.simulate_toys(12, livetime_days=0.02, seed=9)
    ll.set_toy_data(td)
    assert ll.n_toys == 12
    zs, mult = wl.scan_points(12, 2, 2, seed=12, z_range=(-1., 1.))
    params = np.column_stack([mult, zs])
    got = ll.batch_toys(params, names, livetime_days=0.02)
    assert np.all(np.isfinite(got))
    for t in (0, 5, 11):
        ll.set_data(td.to_records(t))
        assert got[t] == ll(livetime_days=0.02, **dict(zip(names, [float(v) for v in params[t]])))
    # a list of record arrays gives the same toys
    ll.set_toy_data([td.to_records(t) for t in range(12)])
    assert np.array_equal(ll.batch_toys(params, names, livetime_days=0.02), got)
