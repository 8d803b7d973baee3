"""Host-side logic of the API mirror that needs no device: parameter bookkeeping and errors, config
handling of Model/Source, histogram semantics, anchor-grid ordering, batch planning, objectives."""
import os
from collections import OrderedDict

import numpy as np
import pytest

import blueice_b200 as bi
from blueice_b200 import engine, pdf_morphers
from blueice_b200.exceptions import (InvalidParameter, InvalidParameterSpecification, NoOpimizationNecessary,
                                     NoShapeParameters, NotPreparedException)
from blueice_b200.hist import Histdd
from blueice_b200.model import Model
from blueice_b200.test_helpers import FixedSampleSource, conf_for_test, make_data
from blueice_b200.utils import arrays_to_grid, combine_dicts, deterministic_hash, InterpolateAndExtrapolate1D
from oracle import hist as ohist
from oracle import morph as omorph


def test_package_exposes_reference_names():
    for name in ['Model', 'Source', 'HistogramPdfSource', 'DensityEstimatingSource', 'MonteCarloSource',
                 'UnbinnedLogLikelihood', 'BinnedLogLikelihood', 'LogLikelihoodSum', 'InvalidParameter',
                 'NotPreparedException', 'OptimizationFailed', 'PDFNotComputedException']:
        assert hasattr(bi, name)
    for method in ['bestfit_scipy', 'one_parameter_interval', 'make_objective', 'best_anchor', 'batch']:
        assert hasattr(bi.UnbinnedLogLikelihood, method)


def test_model_rates_and_source_lookup():
    # mirrors tests/test_model.py of the reference
    m = Model(conf_for_test(n_sources=1))
    np.testing.assert_array_equal(m.expected_events(), np.array([1000]))
    m.sources[0].fraction_in_range = 0.5
    np.testing.assert_array_equal(m.expected_events(), np.array([500]))
    conf = conf_for_test(n_sources=2)
    conf['some_multiplier'] = 2
    m = Model(conf)
    np.testing.assert_array_equal(m.expected_events(), np.array([2000, 2000]))
    assert m.get_source(1) == m.sources[1]
    assert m.get_source_i('s1') == 1 and m.get_source_i(1) == 1
    with pytest.raises(ValueError):
        m.get_source_i('nope')
    conf = conf_for_test(n_sources=1)
    conf['strlen_multiplier'] = 'hi'
    np.testing.assert_array_equal(Model(conf).expected_events(), np.array([2000]))
    with pytest.raises(ValueError):
        Model(conf_for_test(rate_multiplier=3))
    assert Model(conf_for_test(events_per_day=1)).expected_events().dtype == np.float64


def test_per_source_rate_multiplier_setting():
    conf = conf_for_test(n_sources=2, s1_rate_multiplier=3)
    m = Model(conf)
    np.testing.assert_array_equal(m.expected_events(), np.array([1000., 3000.]))
    assert all('s1_rate_multiplier' not in s.config for s in m.sources)


def test_range_cut_is_closed_interval():
    m = Model(conf_for_test())
    d = np.zeros(4, dtype=[('x', float), ('source', int)])
    d['x'] = [-10, 10, -10.0001, 3]
    assert len(m.range_cut(d)) == 3


def test_density_estimating_source_histogram():
    data, n = make_data([dict(n_events=24, x=0.5), dict(n_events=56, x=1.5), dict(n_events=20, x=7.)])
    conf = conf_for_test(events_per_day=42, analysis_space=[['x', [0, 1, 5]]],
                         default_source_class=FixedSampleSource, data=data)
    s = Model(conf).sources[0]
    assert s.fraction_in_range == 0.8
    pmf, n_ev = s.get_pmf_grid()
    np.testing.assert_allclose(pmf, [0.3, 0.7])
    np.testing.assert_array_equal(n_ev, [24, 56])
    assert s.expected_events == 42 * 0.8


def test_histdd_matches_oracle_rules():
    rng = np.random.default_rng(3)
    edges = [np.linspace(0, 1, 6), np.array([0., 0.5, 2., 3.])]
    x = np.r_[edges[0], rng.uniform(-0.2, 1.2, 500)]
    y = np.r_[edges[1], 1.0, 1.0, rng.uniform(-0.5, 3.5, 500)]
    h = Histdd(x, y, bins=edges)
    np.testing.assert_array_equal(h.histogram, ohist.histogramdd(edges, [x, y]))
    h.histogram = rng.random(h.histogram.shape)
    np.testing.assert_array_equal(h.lookup(x, y), ohist.lookup_piecewise(h.histogram, edges, [x, y]))
    np.testing.assert_array_equal(h.bin_centers(0), ohist.bin_centers(edges[0]))
    assert (h * 2.0).histogram[1, 1] == 2 * h.histogram[1, 1]
    np.random.seed(0)
    r = h.get_random(1000)
    assert r.shape == (1000, 2) and r[:, 0].min() >= 0 and r[:, 1].max() <= 3


def test_utils():
    np.testing.assert_array_equal(arrays_to_grid([np.array([1, 2]), np.array([3, 4])]),
                                  np.array([[[1, 3], [1, 4]], [[2, 3], [2, 4]]]))
    assert combine_dicts(dict(a=1, b=2), dict(b=3, c=4), exclude=['c']) == dict(a=1, b=3)
    assert deterministic_hash(dict(a=[1, 2], b=np.arange(3))) == deterministic_hash(dict(b=np.arange(3), a=[1, 2]))
    itp = InterpolateAndExtrapolate1D([0, 1], [0, 42])
    assert itp(3) == 42 and itp(0.5) == 21 and itp([3]) == [42]
    assert InterpolateAndExtrapolate1D(0, 42)(3) == 42


def test_morpher_contract_and_anchor_order():
    with pytest.raises(NoShapeParameters):
        pdf_morphers.GridInterpolator(config={}, shape_parameters=OrderedDict())
    pars = OrderedDict([('a', ({2: 2, -2: -2, 0: 0}, None, None)), ('b', ({1: 'x', 0: 'y'}, None, 0))])
    mr = pdf_morphers.MORPHERS['GridInterpolator']({}, pars)
    pts = mr.get_anchor_points(bounds=None)
    assert isinstance(pts, list) and isinstance(pts[0], tuple)
    assert pts == omorph.anchor_points(omorph.anchor_axes([[2, -2, 0], [1, 0]]))
    tensor = mr.anchor_tensor(lambda z: np.array([z[0] * 10 + z[1]]), [1], {p: p for p in pts})
    assert tensor.shape == (3, 2, 1) and tensor[2, 1, 0] == 21 and tensor[0, 0, 0] == -20


def test_parameter_bookkeeping_and_errors():
    lf = bi.UnbinnedLogLikelihood(conf_for_test(n_sources=2))
    lf.add_rate_parameter('s0')
    with pytest.raises(InvalidParameterSpecification):
        lf.add_shape_parameter('strlen_multiplier', {1: 'x', 2: 'hi', 3: 'wha'})
    with pytest.raises(InvalidParameterSpecification):
        lf.add_shape_parameter('strlen_multiplier', ['x', 'hi'])
    with pytest.raises(InvalidParameterSpecification):
        lf.add_shape_parameter('some_multiplier', (0.5, 1, 2), base_value=1)
    lf.add_shape_parameter('strlen_multiplier', {1: 'q', 2: 'hi', 3: 'wha'}, base_value=1)
    lf.add_shape_parameter('some_multiplier', (0.5, 1, 2, 4))
    assert lf.get_bounds('some_multiplier') == (0.5, 4)
    assert lf.get_bounds('strlen_multiplier') == (1, 3)
    assert lf.get_bounds() == [(1, 3), (0.5, 4)]
    assert lf.get_bounds('s0_rate_multiplier') == (0, float('inf'))
    with pytest.raises(InvalidParameter):
        lf.get_bounds('blargh')
    mult, settings = lf._kwargs_to_settings(s1_rate_multiplier=2.5, some_multiplier=2)
    assert mult == [1, 2.5] and settings == dict(strlen_multiplier=1, some_multiplier=2)
    with pytest.raises(InvalidParameter):
        lf._kwargs_to_settings(blargh=41)
    with pytest.raises(ValueError):
        lf._kwargs_to_settings(strlen_multiplier='hi')
    with pytest.raises(ValueError):
        lf._kwargs_to_settings(some_multiplier=np.int64(2))      # reference quirk: np.int64 is not accepted
    assert lf.parameter_names() == ['s0_rate_multiplier', 'strlen_multiplier', 'some_multiplier']
    d = np.zeros(3, dtype=[('x', float), ('source', int)])
    with pytest.raises(NotPreparedException):
        lf.set_data(d)
    with pytest.raises(NotPreparedException):
        lf()


def test_allow_negative_bounds():
    conf = conf_for_test(n_sources=2)
    conf['sources'][1]['allow_negative'] = True
    lf = bi.UnbinnedLogLikelihood(conf)
    assert lf.get_bounds('s1_rate_multiplier') == (float('-inf'), float('inf'))
    assert lf.get_bounds('s0_rate_multiplier') == (0, float('inf'))
    assert lf.source_allowed_negative == [False, True]


def test_make_objective_names_bounds_guesses():
    lf = bi.UnbinnedLogLikelihood(conf_for_test(n_sources=2))
    lf.add_rate_parameter('s0')
    lf.add_rate_parameter('s1')
    lf.add_shape_parameter('some_multiplier', (0.5, 1, 2, 4))
    f, names, guess, bounds = lf.make_objective(s1_rate_multiplier=2.)
    assert names == ['s0_rate_multiplier', 'some_multiplier']
    np.testing.assert_array_equal(guess, [1, 1])
    assert bounds == [(0, None), (0.5, 4)]
    f, names, guess, bounds = lf.make_objective(rates_in_log_space=True, guess=dict(s0_rate_multiplier=10.))
    assert guess[0] == 1.0 and bounds[0] == (None, None)
    with pytest.raises(NoOpimizationNecessary):
        lf.make_objective(s0_rate_multiplier=1, s1_rate_multiplier=1, some_multiplier=1)


def test_grid_cell_ids_follow_the_oracle_rule():
    rng = np.random.default_rng(0)
    axes = [np.array([-2., -1., 0., 1., 2.]), np.array([0.5, 1., 4.]), np.array([3.])]
    grid = engine.MorphGrid(axes)
    zs = np.column_stack([rng.uniform(-2, 2, 500), rng.uniform(0.5, 4, 500), np.full(500, 3.)])
    zs[:5, 0] = axes[0]
    zs[:3, 1] = axes[1]
    ids = grid.cell_ids(zs)
    for row, cid in zip(zs, ids):
        c0 = omorph.find_cell(axes[0], row[0])[0]
        c1 = omorph.find_cell(axes[1], row[1])[0]
        assert cid == (c0 * 2 + c1) * 1 + 0
    ok = grid.in_range(np.array([[0., 1., 3.], [2.5, 1., 3.], [np.nan, 1., 3.], [0., 1., 3.1]]))
    assert ok.tolist() == [True, False, False, False]
    assert grid.n_anchors == 15 and grid.n_corners == 8


def test_plan_points_covers_every_in_range_point_exactly_once():
    rng = np.random.default_rng(1)
    grid = engine.MorphGrid([np.array([-2., -1., 0., 1., 2.])] * 2)
    zs = rng.uniform(-2.3, 2.0, (3000, 2))
    zs[7] = np.nan
    plan = engine.plan_points(grid, zs)
    covered = np.zeros(len(zs), dtype=int)
    covered[plan.stream_points] += 1
    assert np.array_equal(covered == 1, plan.in_range)
    assert not plan.in_range[7]
    assert np.all(np.diff(plan.stream_points) > 0)          # in batch order


def test_group_pairs_invariants():
    """Host-side schedule of the template-space kernels: every pair in exactly one group, groups never larger than
    allowed, a group never mixes datasets or hypercube cells, out-of-range pairs stay alone."""
    from blueice_b200.engine import group_pairs
    rng = np.random.default_rng(0)
    for trial in range(20):
        Q = int(rng.integers(0, 400))
        n_cells = int(rng.integers(1, 20))
        dataset = rng.integers(0, 7, size=Q)
        cells = rng.integers(-1, n_cells, size=Q)
        gp = int(rng.choice([1, 4, 8]))
        order, first, count = group_pairs(dataset, cells, n_cells, gp)
        assert sorted(order.tolist()) == list(range(Q))
        assert int(count.sum()) == Q and (Q == 0 or (count.min() >= 1 and count.max() <= gp))
        covered = np.zeros(Q, dtype=int)
        for f, c in zip(first, count):
            members = order[f:f + c]
            covered[members] += 1
            assert len(set(dataset[members].tolist())) == 1 and len(set(cells[members].tolist())) == 1
            if cells[members[0]] < 0:
                assert c == 1
        assert np.all(covered == 1)
        # groups are laid out back to back in sorted order
        assert Q == 0 or (first[0] == 0 and np.array_equal(first[1:], np.cumsum(count)[:-1]))
    # stable: pairs of one (dataset, cell) keep their relative order
    order, first, count = group_pairs([1, 0, 1, 0, 1], [2, 2, 2, 2, 2], 5, 8)
    assert order.tolist() == [1, 3, 0, 2, 4] and first.tolist() == [0, 2] and count.tolist() == [2, 3]


def test_mixture_group_width_rules():
    """Points per group of a mixture-form (K5b) schedule: 1 point streams alone, up to 8 points share one pass over
    the events, and 16 (two m-tiles per warp on the FP64 tensor pipe) only for more than 8 points on datasets that
    outgrow the L2; the gather kernel (BI_MIX_MMA=0) never gets wide groups.  group_pairs cuts runs accordingly."""
    from blueice_b200 import _cabi
    from blueice_b200.engine import _MIX_WIDE_MIN_SUPERBLOCKS, group_pairs, mixture_group_width
    big, small = _MIX_WIDE_MIN_SUPERBLOCKS, _MIX_WIDE_MIN_SUPERBLOCKS - 1
    assert mixture_group_width(0, big) == 1 and mixture_group_width(1, big) == 1
    assert mixture_group_width(2, big) == 8 and mixture_group_width(8, big) == 8
    assert mixture_group_width(9, big) == 16 and mixture_group_width(4096, big) == 16
    assert mixture_group_width(11, small) == 8
    assert mixture_group_width(11, big, tensor_pipe=False) == 8
    assert mixture_group_width(11, 3, wide_min_superblocks=0) == 16          # the override the GPU tests use
    assert _cabi.MIX_GROUP_POINTS == 8 and _cabi.MIX_GROUP_POINTS_WIDE == 16
    # an 11-point finite-difference batch in one cell: one wide group, or 8 + 3
    cells = np.zeros(11, dtype=np.int64)
    for width, want in ((16, [11]), (8, [8, 3])):
        _, first, count = group_pairs(np.zeros(11, dtype=np.int64), cells, 1, width)
        assert count.tolist() == want and first.tolist() == np.concatenate([[0], np.cumsum(want)[:-1]]).tolist()


def test_toy_tables_and_means_follow_model_simulate():
    """Host side of on-device toy generation: the cdf tables are Histdd.get_random's (cumsum of the flattened pmf, last
    entry 1) and the Poisson means are Model.simulate's (model.py:80-84: rate multipliers, livetime, fraction in range)."""
    from blueice_b200 import HistogramPdfSource, Model
    from blueice_b200 import toys
    from blueice_b200.hist import Histdd
    from blueice_b200.test_helpers import conf_for_test

    edges = [np.linspace(0, 10, 6), np.linspace(-1, 1, 4)]

    class Flat(HistogramPdfSource):
        def build_histogram(self):
            dens = np.arange(1., 16.).reshape(5, 3) * self.config.get('tilt', 1.0)
            vol = np.outer(np.diff(edges[0]), np.diff(edges[1]))
            dens = dens / np.sum(dens * vol)
            self._pdf_histogram = Histdd.from_histogram(dens, edges, axis_names=['a', 'b'])
            self._bin_volumes = self._pdf_histogram.bin_volumes()
            self._n_events_histogram = Histdd.from_histogram(np.full(dens.shape, np.inf), edges, ['a', 'b'])

    config = dict(sources=[dict(name='s0', events_per_day=30.), dict(name='s1', events_per_day=5., tilt=2.0)],
                  default_source_class=Flat, analysis_space=[['a', edges[0]], ['b', edges[1]]], livetime_days=2.,
                  force_recalculation=True, never_save_to_cache=True)
    model = Model(config)
    got_edges, cdf = toys.source_tables(model)
    assert all(np.array_equal(a, b) for a, b in zip(got_edges, edges)) and cdf.shape == (2, 15)
    assert np.all(np.diff(cdf, axis=1) > 0) and np.allclose(cdf[:, -1], 1.0)
    pmf0 = (model.sources[0]._pdf_histogram.histogram * model.sources[0]._bin_volumes).ravel()
    assert np.array_equal(cdf[0], np.cumsum(pmf0) / np.cumsum(pmf0)[-1])
    mus = toys.toy_means(model, rate_multipliers={'s1': 3.0}, livetime_days=1.0)
    expect = [model.expected_events(model.sources[0]) * 1 / model.sources[0].fraction_in_range * 0.5,
              model.expected_events(model.sources[1]) * 3.0 / model.sources[1].fraction_in_range * 0.5]
    assert np.array_equal(mus, np.asarray(expect))
    # sources that are not stock histogram templates cannot be generated on the device
    gauss = Model(conf_for_test(n_sources=1))
    with pytest.raises(NotImplementedError):
        toys.source_tables(gauss)


def test_install_as_blueice_aliases_the_reference_import_paths():
    """Code written against the reference imports `blueice.likelihood` etc.; the alias makes those names this package."""
    import subprocess
    import sys
    code = ("import blueice_b200, sys\n"
            "blueice_b200.install_as_blueice()\n"
            "import blueice\n"
            "from blueice.likelihood import UnbinnedLogLikelihood, BinnedLogLikelihood, LogLikelihoodSum\n"
            "from blueice.source import HistogramPdfSource, DensityEstimatingSource\n"
            "from blueice.model import Model\n"
            "from blueice.inference import bestfit_scipy, one_parameter_interval\n"
            "from blueice.exceptions import InvalidParameter, NotPreparedException\n"
            "from blueice.pdf_morphers import MORPHERS\n"
            "import blueice.likelihood, blueice_b200.likelihood\n"
            "assert blueice.likelihood is blueice_b200.likelihood and blueice is blueice_b200\n"
            "assert UnbinnedLogLikelihood is blueice_b200.UnbinnedLogLikelihood\n"
            "try:\n"
            "    import blueice.parallel\n"
            "except ImportError:\n"
            "    pass\n"
            "else:\n"
            "    raise SystemExit('blueice.parallel should not resolve')\n"
            "print('ok')\n")
    env = dict(os.environ)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env["PYTHONPATH"] = root                       # the reference is NOT on the path: every name must come from the alias
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=root)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr


def test_capture_graph_keeps_the_cyclic_collector_off_during_a_capture():
    """engine.capture_graph: a collection inside a CUDA graph capture runs destructors that call CUDA APIs a capturing thread
    must not call (the process aborts); the collector is off inside the capture and restored afterwards, also on errors."""
    import gc
    seen = []

    class FakeGraphContext(object):
        def __init__(self, graph, capture_error_mode=None):
            seen.append(('mode', capture_error_mode))

        def __enter__(self):
            seen.append(('inside', gc.isenabled()))

        def __exit__(self, *exc):
            return False

    class FakeTorch(object):
        class cuda(object):
            graph = FakeGraphContext

    assert gc.isenabled()
    with engine.capture_graph(FakeTorch, object()):
        assert not gc.isenabled()
    assert gc.isenabled()
    with pytest.raises(RuntimeError):
        with engine.capture_graph(FakeTorch, object()):
            raise RuntimeError("capture failed")
    assert gc.isenabled()
    gc.disable()
    try:
        with engine.capture_graph(FakeTorch, object()):
            pass
        assert not gc.isenabled()                    # a caller that runs with the collector off keeps it off
    finally:
        gc.enable()
    assert ('mode', 'thread_local') in seen and ('inside', False) in seen
