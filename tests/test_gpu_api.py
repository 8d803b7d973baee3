"""API-level parity: the blueice_b200 classes replay (a) golden vectors produced by the unmodified
reference (tests/golden/*.npz) and (b) the reference's own test scenarios
(tests/test_likelihood.py, test_binned_likelihood.py, test_BeestonBarlow.py, test_morphers.py,
test_inference.py, test_source.py of JelleAalbers/blueice v1.2.1), with the same fixtures.

Exact `==` asserts of the reference that compare two evaluations of the SAME likelihood are kept exact.
Asserts that compare against a scipy closed form (which needs device log() to round like NumPy's) use
the north-star tolerance |dlogL| <= 1e-9 * N_events, as SURVEY.md section 7 prescribes.
"""
from collections import OrderedDict

import numpy as np
import pytest
import scipy.stats as sps
from scipy import stats

import bench_workloads as wl
from conftest import load_golden

pytestmark = pytest.mark.gpu

TOL = 1e-9


def close(a, b, n_events=1):
    return abs(a - b) <= TOL * max(n_events, 1)


def assert_vector_close(got, ref, n_events):
    got, ref = np.asarray(got, dtype=float), np.asarray(ref, dtype=float)
    assert np.array_equal(np.isneginf(got), np.isneginf(ref))
    fin = np.isfinite(ref)
    assert np.all(np.abs(got[fin] - ref[fin]) <= TOL * n_events), np.abs(got[fin] - ref[fin]).max()
    assert np.all(np.abs(got[fin] - ref[fin]) <= 2e-13 * (np.abs(ref[fin]) + n_events)), np.abs(got[fin] - ref[fin]).max()


# ------------------------------------------------------------------------------------------------
# (a) golden vectors from the unmodified reference
# ------------------------------------------------------------------------------------------------
def test_golden_config1_gaussian():
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.test_helpers import conf_for_test
    g = load_golden('c1_gaussian')
    lf = UnbinnedLogLikelihood(conf_for_test(n_sources=1))
    lf.add_rate_parameter('s0')
    lf.add_shape_parameter('mu', {-2: -2, 0: 0, 2: 2})
    lf.prepare()
    d = np.zeros(len(g['x']), dtype=[('x', float), ('source', int)])
    d['x'] = g['x']
    lf.set_data(d)
    n = len(d)
    scalar = np.array([lf(mu=float(z[0]), s0_rate_multiplier=float(m[0])) for z, m in zip(g['zs'], g['mult'])])
    assert_vector_close(scalar, g['logl'], n)
    batch = lf.batch(np.column_stack([g['mult'], g['zs']]), ['s0_rate_multiplier', 'mu'])
    assert np.array_equal(batch, scalar)                              # batch rows == scalar calls, bitwise
    assert np.array_equal(lf.batch(np.column_stack([g['mult'], g['zs']])), scalar)   # default column order
    for k, i in enumerate(g['full_index']):
        ll, mus, ps = lf(mu=float(g['zs'][i, 0]), s0_rate_multiplier=float(g['mult'][i, 0]), full_output=True)
        assert ll == scalar[i]
        assert np.array_equal(mus, g['full_mus'][k])                  # bit-exact morph of the rates
        assert np.array_equal(ps, g['full_ps'][k])                    # bit-exact morph of the pdf values


@pytest.mark.parametrize("method", ["linear", "piecewise"])
def test_golden_config2_templates(method):
    from blueice_b200 import HistogramPdfSource, UnbinnedLogLikelihood
    from blueice_b200.hist import Histdd
    g = load_golden('c2_templates_' + method)
    n_sources, n_shape, anchors, bins = 2, 2, (-1., 0., 1.), (20, 16)
    axes, edges, templates, mus = wl.c2_arrays(n_sources, n_shape, anchors, bins)
    names, params = ['cs1', 'cs2'], ['shift1', 'shift2']
    cls = wl.array_source_class(HistogramPdfSource, Histdd, axes, edges, names, mus, templates, None, params)
    lf = UnbinnedLogLikelihood(wl.array_model_config(cls, edges, names, n_sources, params, method))
    for s in range(n_sources):
        lf.add_rate_parameter('src%d' % s)
    for p in params:
        lf.add_shape_parameter(p, anchors)
    lf.prepare()
    d = np.zeros(len(g['x']), dtype=[('cs1', float), ('cs2', float), ('source', int)])
    d['cs1'], d['cs2'] = g['x'], g['y']
    lf.set_data(d)
    # the anchor tensor built by the K3 gather equals the reference's, bit for bit
    dev = lf._engine.ps_anchor[:, :, :len(d)].cpu().numpy().reshape(3, 3, 2, len(d))
    assert np.array_equal(dev, g['ps_anchor'])
    table = np.column_stack([g['mult'], g['zs']])
    got = lf.batch(table, ['src0_rate_multiplier', 'src1_rate_multiplier', 'shift1', 'shift2'])
    assert_vector_close(got, g['logl'], len(d))
    for k, i in enumerate(g['full_index']):
        ll, mus_i, ps_i = lf(shift1=float(g['zs'][i, 0]), shift2=float(g['zs'][i, 1]), full_output=True,
                             src0_rate_multiplier=float(g['mult'][i, 0]), src1_rate_multiplier=float(g['mult'][i, 1]))
        assert ll == got[i]
        assert np.array_equal(mus_i, g['full_mus'][k]) and np.array_equal(ps_i, g['full_ps'][k])
    # the public per-source pdf call goes through the same kernel
    src = lf.anchor_models[(0., 1.)].sources[1]
    assert np.array_equal(src.pdf(g['x'], g['y']), g['ps_anchor'][1, 2, 1])


@pytest.mark.parametrize("bb", [False, True])
def test_golden_binned(bb):
    from blueice_b200 import BinnedLogLikelihood, HistogramPdfSource
    from blueice_b200.hist import Histdd
    g = load_golden('binned_bb' if bb else 'binned_plain')
    bins, n_sources, n_shape, anchors = (6, 5, 4), 3, 2, (-1., 0., 1.)
    axes, edges, mus, pmf, n_model, observed = wl.c3_arrays(bins, n_sources, n_shape, anchors, seed=3, total_events=600.)
    vol = np.ones(1)
    for e in edges:
        vol = np.multiply.outer(vol, np.diff(e))
    density = pmf / vol.reshape(bins)
    names, params = ['x', 'y', 'z'], ['za', 'zb']
    cls = wl.array_source_class(HistogramPdfSource, Histdd, axes, edges, names, mus, density, n_model, params)
    cfg = {'model_statistical_uncertainty_handling': 'bb_single', 'bb_single_source': 0} if bb else None
    lf = BinnedLogLikelihood(wl.array_model_config(cls, edges, names, n_sources, params), cfg)
    for s in range(n_sources):
        lf.add_rate_parameter('src%d' % s)
    for p in params:
        lf.add_shape_parameter(p, anchors)
    lf.prepare()
    d = np.zeros(len(g['x']), dtype=[('x', float), ('y', float), ('z', float), ('source', int)])
    d['x'], d['y'], d['z'] = g['x'], g['y'], g['z']
    lf.set_data(d)
    assert np.array_equal(lf.data_events_per_bin.histogram, g['observed'])        # device binning, exact
    col_names = ['src%d_rate_multiplier' % s for s in range(n_sources)] + params
    table = np.column_stack([g['mult'], g['zs']])
    ok = ~g['raises']
    got = lf.batch(table[ok], col_names)
    ref = g['logl'][ok]
    assert np.array_equal(np.isneginf(got), np.isneginf(ref))
    fin = np.isfinite(ref)
    assert np.all(np.abs(got[fin] - ref[fin]) <= TOL * len(d))
    assert np.all(np.abs(got[fin] - ref[fin]) <= 1e-12 * (np.abs(ref[fin]) + 120))
    for i in np.flatnonzero(g['raises']):                                          # reference asserts -> so do we
        with pytest.raises(AssertionError):
            lf(**dict(zip(col_names, [float(v) for v in table[i]])))
    for k, i in enumerate(g['full_index']):
        ll, mus_i, pmfs_i = lf(full_output=True, **dict(zip(col_names, [float(v) for v in table[i]])))
        np.testing.assert_allclose(mus_i, g['full_mus'][k], rtol=1e-12)
        np.testing.assert_allclose(pmfs_i, g['full_pmfs'][k], rtol=1e-11)
        if not bb:
            assert np.array_equal(pmfs_i, g['full_pmfs'][k]) and np.array_equal(mus_i, g['full_mus'][k])


def test_golden_multisource():
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.test_helpers import conf_for_test
    g = load_golden('multisource')
    lf = UnbinnedLogLikelihood(conf_for_test(n_sources=2))
    lf.add_shape_parameter('some_multiplier', (0.5, 1, 2, 4))
    lf.add_rate_parameter('s0')
    lf.add_rate_parameter('s1')
    lf.prepare()
    d = np.zeros(len(g['x']), dtype=[('x', float), ('source', int)])
    d['x'] = g['x']
    lf.set_data(d)
    names = ['s0_rate_multiplier', 's1_rate_multiplier', 'some_multiplier']
    assert_vector_close(lf.batch(g['params'], names), g['logl'], len(d))


def test_golden_long_contraction():
    """3 shape parameters x 5 sources = 40 contraction terms (the K-chunk form of K2) against values of the unmodified
    reference (tests/golden/make_golden.py: golden_long_contraction)."""
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.test_helpers import conf_for_test
    g = load_golden('long_contraction')
    lf = UnbinnedLogLikelihood(conf_for_test(n_sources=5, events_per_day=300.))
    lf.add_shape_parameter('mu', (-0.5, 0., 0.5))
    lf.add_shape_parameter('sigma', (0.8, 1., 1.3))
    lf.add_shape_parameter('some_multiplier', (0.5, 1., 2.))
    for i in range(5):
        lf.add_rate_parameter('s%d' % i)
    lf.prepare()
    d = np.zeros(len(g['x']), dtype=[('x', float), ('source', int)])
    d['x'] = g['x']
    lf.set_data(d)
    assert lf._engine.n_terms == 40
    names = [str(v) for v in g['names']]
    got = lf.batch(g['params'], names)
    assert np.isneginf(g['logl']).sum() == 2
    assert_vector_close(got, g['logl'], len(d))
    for i in (0, 1, 5, 17):                                  # scalar calls: the same bits as the batch rows
        assert lf(**dict(zip(names, [float(v) for v in g['params'][i]]))) == got[i]


def test_golden_sourcewise():
    """source_wise_interpolation (likelihood.py:113-145,210-240,534-555): s0 on (mu, sigma), s1 on sigma, s2 on nothing."""
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.test_helpers import conf_for_test
    g = load_golden('sourcewise')
    config = conf_for_test(n_sources=3)
    config['sources'][0]['events_per_day'] = 700.
    config['sources'][1].update(events_per_day=250., extra_dont_hash_settings=['mu'])
    config['sources'][2].update(events_per_day=50., extra_dont_hash_settings=['mu', 'sigma'])
    config['source_wise_interpolation'] = True
    lf = UnbinnedLogLikelihood(config)
    for s in range(3):
        lf.add_rate_parameter('s%d' % s)
    lf.add_shape_parameter('mu', {-2: -2, 0: 0, 2: 2})
    lf.add_shape_parameter('sigma', (0.5, 1, 2))
    lf.prepare()
    assert list(lf.source_shape_parameters.keys()) == ['s0', 's1']
    assert lf._get_shape_indices('s1') == [1] and lf._get_model_anchor((2,), 's1') == (None, 2)
    d = np.zeros(len(g['x']), dtype=[('x', float), ('source', int)])
    d['x'] = g['x']
    lf.set_data(d)
    n = len(d)
    eng = lf._engine
    assert eng.n_rows == 13 and eng.n_terms == 4 + 2 + 1
    # the per-(source, sub-anchor) rows equal the reference's own, bit for bit
    assert np.array_equal(eng.ps_anchor[:, :n].cpu().numpy(), g['rows'])
    assert np.array_equal(eng.mus_rows_host, g['mus_rows'])
    names = ['s0_rate_multiplier', 's1_rate_multiplier', 's2_rate_multiplier', 'mu', 'sigma']
    table = np.column_stack([g['mult'], g['zs']])
    batch = lf.batch(table, names)
    assert_vector_close(batch, g['logl'], n)
    for i in (0, 5, 17, 39):
        assert lf(**dict(zip(names, [float(v) for v in table[i]]))) == batch[i]
    for k, i in enumerate(g['full_index']):
        ll, mus, ps = lf(full_output=True, **dict(zip(names, [float(v) for v in table[i]])))
        assert ll == batch[i]
        assert np.array_equal(mus, g['full_mus'][k])
        assert np.array_equal(ps, g['full_ps'][k])
    np.testing.assert_allclose(lf.mus_interpolator(g['zs'][30]) * g['mult'][30], g['full_mus'][0], rtol=1e-15)


def test_source_wise_interpolation():
    """tests/test_likelihood.py:95-121: identical results with and without source-wise interpolation."""
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.test_helpers import conf_for_test
    data = np.zeros(5, dtype=[('x', float), ('source', int)])
    data['x'] = np.linspace(0, 1, 5)
    config = conf_for_test(events_per_day=1)
    lf = UnbinnedLogLikelihood(config)
    lf.add_shape_parameter("mu", anchors={-2: -2, 0: 0, 2: 2})
    lf.prepare()
    lf.set_data(data)
    ret_0 = lf(full_output=True)
    ret_1 = lf(full_output=True, mu=1)
    config["source_wise_interpolation"] = True
    lf_source_wise = UnbinnedLogLikelihood(config)
    lf_source_wise.add_shape_parameter("mu", anchors={-2: -2, 0: 0, 2: 2})
    lf_source_wise.prepare()
    lf_source_wise.set_data(data)
    ret_source_wise_0 = lf_source_wise(full_output=True)
    ret_source_wise_1 = lf_source_wise(full_output=True, mu=1)
    assert ret_0[0] == ret_source_wise_0[0]
    assert (ret_0[1] == ret_source_wise_0[1]).all()
    assert (ret_0[2] == ret_source_wise_0[2]).all()
    assert ret_1[0] == ret_source_wise_1[0]
    assert (ret_1[1] == ret_source_wise_1[1]).all()
    assert (ret_1[2] == ret_source_wise_1[2]).all()
    from blueice_b200 import BinnedLogLikelihood
    lb = BinnedLogLikelihood(config)
    lb.add_shape_parameter("mu", anchors={-2: -2, 0: 0, 2: 2})
    with pytest.raises(NotImplementedError):
        lb.prepare()


# ------------------------------------------------------------------------------------------------
# (b) the reference's own test scenarios
# ------------------------------------------------------------------------------------------------
def test_likelihood_value():
    """tests/test_likelihood.py:8-18."""
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.test_helpers import conf_for_test
    lf = UnbinnedLogLikelihood(conf_for_test(events_per_day=1))
    lf.add_rate_parameter('s0')
    lf.set_data(np.zeros(1, dtype=[('x', float), ('source', int)]))
    assert close(lf(), -1 + stats.norm.logpdf(0))
    assert close(lf(s0_rate_multiplier=2), -2 + np.log(2 * stats.norm.pdf(0)))
    # deliberate deviation: float rates, no integer truncation of mus
    assert close(lf(s0_rate_multiplier=2.5), -2.5 + np.log(2.5 * stats.norm.pdf(0)))


def test_no_shape_params():
    """tests/test_likelihood.py:21-35."""
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.test_helpers import conf_for_test
    lf = UnbinnedLogLikelihood(conf_for_test())
    d = lf.base_model.simulate()
    lf.prepare()
    lf.set_data(d)
    assert np.isfinite(lf())
    lf = UnbinnedLogLikelihood(conf_for_test(mc=True, n_events_for_pdf=int(1e5)))
    d = lf.base_model.simulate()
    lf.prepare()
    lf.set_data(d)
    assert np.isfinite(lf())


def test_shape_params():
    """tests/test_likelihood.py:36-58."""
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.exceptions import InvalidParameterSpecification
    from blueice_b200.test_helpers import conf_for_test
    lf = UnbinnedLogLikelihood(conf_for_test(n_sources=1))
    lf.add_rate_parameter('s0')
    with pytest.raises(InvalidParameterSpecification):
        lf.add_shape_parameter('strlen_multiplier', {1: 'x', 2: 'hi', 3: 'wha'})
    lf.add_shape_parameter('strlen_multiplier', {1: 'q', 2: 'hi', 3: 'wha'}, base_value=1)
    d = lf.base_model.simulate()
    lf.prepare()
    lf.set_data(d)
    assert len(lf.anchor_models) == 3
    with pytest.raises(ValueError):
        lf(strlen_multiplier='hi')
    lf(strlen_multiplier=1.5)
    assert lf() == lf(strlen_multiplier=1)
    assert lf(strlen_multiplier=1.5) < lf()


def test_rate_and_shape_uncertainty_priors():
    """tests/test_likelihood.py:61-92."""
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.exceptions import InvalidParameterSpecification
    from blueice_b200.test_helpers import conf_for_test
    one = np.zeros(1, dtype=[('x', float), ('source', int)])
    log_prior = stats.norm(1, 0.5).logpdf
    lf = UnbinnedLogLikelihood(conf_for_test(events_per_day=1))
    lf.add_rate_uncertainty('s0', 0.5)
    lf.set_data(one)
    assert close(lf(), -1 + stats.norm.logpdf(0) + log_prior(1))
    assert close(lf(s0_rate_multiplier=2), -2 + np.log(2 * stats.norm.pdf(0)) + log_prior(2))

    lf = UnbinnedLogLikelihood(conf_for_test(events_per_day=1))
    with pytest.raises(InvalidParameterSpecification):
        lf.add_shape_uncertainty('strlen_multiplier', 0.5, {1: 'x', 2: 'hi', 3: 'wha'})
    lf.add_shape_uncertainty(setting_name='strlen_multiplier', fractional_uncertainty=0.5,
                             anchor_zs={1: 'x', 2: 'hi', 3: 'wha'}, base_value=1)
    lf.prepare()
    lf.set_data(one)
    assert close(lf(), -1 + stats.norm.logpdf(0) + log_prior(1))
    assert close(lf(strlen_multiplier=2), -2 + np.log(2 * stats.norm.pdf(0)) + log_prior(2))
    # priors in batch form: same numbers as the scalar calls
    assert np.array_equal(lf.batch(np.array([[1.], [2.], [2.5]]), ['strlen_multiplier']),
                          [lf(strlen_multiplier=1.), lf(strlen_multiplier=2.), lf(strlen_multiplier=2.5)])


def test_multisource_likelihood():
    """tests/test_likelihood.py:124-148 (exact equalities kept exact)."""
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.test_helpers import almost_equal, conf_for_test
    lf = UnbinnedLogLikelihood(conf_for_test(n_sources=2))
    lf.add_shape_parameter('some_multiplier', (0.5, 1, 2, 4))
    lf.add_rate_parameter('s0')
    lf.add_rate_parameter('s1')
    lf.prepare()
    lf.set_data(lf.base_model.simulate())
    assert lf(s0_rate_multiplier=1, s1_rate_multiplier=1, some_multiplier=1) == lf()
    assert lf(s0_rate_multiplier=1, s1_rate_multiplier=1) == lf()
    assert lf(s0_rate_multiplier=1) == lf()
    assert lf(some_multiplier=1) == lf()
    assert almost_equal(lf(s0_rate_multiplier=2), lf(s1_rate_multiplier=2))
    assert almost_equal(lf(s0_rate_multiplier=4), lf(s0_rate_multiplier=2.5, s1_rate_multiplier=2.5))
    assert lf(s0_rate_multiplier=2, s1_rate_multiplier=2) == lf(some_multiplier=2)   # anchor hit: weights {0, 1}
    assert lf(some_multiplier=2) < lf()


def test_error_handling():
    """tests/test_likelihood.py:151-171."""
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.exceptions import InvalidParameter, NotPreparedException
    from blueice_b200.test_helpers import conf_for_test
    lf = UnbinnedLogLikelihood(conf_for_test())
    d = lf.base_model.simulate()
    lf.add_shape_parameter('some_multiplier', (0.5, 1, 2))
    with pytest.raises(NotPreparedException):
        lf.set_data(d)
    with pytest.raises(NotPreparedException):
        lf()
    lf.prepare()
    with pytest.raises(NotPreparedException):
        lf()
    lf.set_data(d)
    lf()
    with pytest.raises(InvalidParameter):
        lf(blargh=41)
    with pytest.raises(InvalidParameter):
        lf.batch(np.ones((2, 1)), ['blargh'])


def test_noninterpolated_pdf():
    """tests/test_likelihood.py:174-188: compute_pdf=True builds a new model at the requested point."""
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.test_helpers import almost_equal, conf_for_test
    conf = conf_for_test(n_sources=1)
    conf['some_multiplier'] = 3e-3
    lf = UnbinnedLogLikelihood(conf)
    lf.add_shape_parameter('mu', (0., 1.))
    lf.add_shape_parameter('sigma', (1., 2.))
    lf.prepare()
    lf.set_data(np.zeros(1, dtype=[('x', float)]))
    target = sps.poisson(3).logpmf(1) + sps.norm(0.5, 1.5).logpdf(0)
    assert almost_equal(lf(compute_pdf=True, mu=0.5, sigma=1.5), target, 1e-5)
    assert not almost_equal(lf(compute_pdf=False, mu=0.5, sigma=1.5), target, 1e-5)


def test_livetime_scaling():
    """tests/test_likelihood.py:204-235."""
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.test_helpers import conf_for_test
    conf = conf_for_test()
    d = np.zeros(1, dtype=[('x', float)])
    lf = UnbinnedLogLikelihood(conf)
    lf.prepare()
    lf.set_data(d)
    orig = lf()
    with pytest.raises(ValueError):
        lf(livetime_days=1)
    conf['livetime_days'] = 1
    lf = UnbinnedLogLikelihood(conf)
    lf.add_rate_parameter('s0')
    lf.prepare()
    lf.set_data(d)
    assert lf(livetime_days=1) == orig
    assert lf(livetime_days=2) == lf(s0_rate_multiplier=2)
    assert lf(livetime_days=0) == lf(s0_rate_multiplier=0)
    conf['livetime_days'] = 0
    lf_zero = UnbinnedLogLikelihood(conf)
    lf_zero.prepare()
    lf_zero.set_data(d)
    with pytest.raises(ValueError):
        lf_zero(livetime_days=1)
    assert lf_zero() == lf(s0_rate_multiplier=0)


def test_unphysical_rates_and_options():
    """Options the reference leaves untested (SURVEY.md section 4 gaps): -inf vs 'error', allow_negative,
    apply_efficiency, outlier_likelihood."""
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.test_helpers import conf_for_test
    d = np.zeros(5, dtype=[('x', float), ('source', int)])
    d['x'] = np.linspace(0, 1, 5)
    lf = UnbinnedLogLikelihood(conf_for_test(n_sources=2))
    lf.add_rate_parameter('s0')
    lf.add_rate_parameter('s1')
    lf.set_data(d)
    assert lf(s0_rate_multiplier=-1) == -np.inf
    assert lf(s0_rate_multiplier=float('inf')) == -np.inf
    assert np.isfinite(lf(s0_rate_multiplier=0))
    assert np.array_equal(lf.batch(np.array([[-1., 1.], [1., 1.]]))[[0]], [-np.inf])
    lf_err = UnbinnedLogLikelihood(conf_for_test(n_sources=2), {'unphysical_behaviour': 'error'})
    lf_err.add_rate_parameter('s0')
    lf_err.set_data(d)
    with pytest.raises(ValueError):
        lf_err(s0_rate_multiplier=-1)
    # a source that may go negative
    conf = conf_for_test(n_sources=2)
    conf['sources'][1]['allow_negative'] = True
    lf_neg = UnbinnedLogLikelihood(conf)
    lf_neg.add_rate_parameter('s0')
    lf_neg.add_rate_parameter('s1')
    lf_neg.set_data(d)
    assert np.isfinite(lf_neg(s1_rate_multiplier=-0.5))
    assert lf_neg(s0_rate_multiplier=-0.5) == -np.inf
    assert lf_neg(s1_rate_multiplier=-1.5) == -np.inf          # total rate negative
    # efficiency as a shape parameter that scales one source's rate
    conf = conf_for_test(n_sources=2, efficiency=0.8)
    conf['sources'][0]['apply_efficiency'] = True
    lf_eff = UnbinnedLogLikelihood(conf)
    lf_eff.add_rate_parameter('s0')
    lf_eff.add_shape_parameter('efficiency', (0.5, 0.8, 1.0))
    lf_eff.prepare()
    lf_eff.set_data(d)
    ll, mus, _ = lf_eff(efficiency=0.6, full_output=True)
    assert np.isclose(mus[0], 600., rtol=1e-14) and np.isclose(mus[1], 1000., rtol=1e-14)
    assert close(lf_eff(efficiency=0.6), lf_eff(efficiency=1.0, s0_rate_multiplier=0.6), 5)
    # outlier likelihood: events far away from the only source
    far = np.zeros(3, dtype=[('x', float), ('source', int)])
    far['x'] = 9.99
    lf_out = UnbinnedLogLikelihood(conf_for_test(sigma=0.01), {'outlier_likelihood': 1e-5})
    lf_out.set_data(far)
    assert close(lf_out(), -1000. + 3 * np.log(1e-5), 3)


def test_single_bin_and_two_bin_binned():
    """tests/test_binned_likelihood.py:10-38 and tests/test_likelihood.py:191-201."""
    from blueice_b200 import BinnedLogLikelihood
    from blueice_b200.test_helpers import almost_equal, conf_for_test
    small = dict(n_events_for_pdf=int(1e5))
    lf = BinnedLogLikelihood(conf_for_test(mc=True, analysis_space=[['x', [-40, 40]]], **small))
    lf.add_rate_parameter('s0')
    lf.prepare()
    lf.set_data(np.zeros(1, dtype=[('x', float), ('source', int)]))
    assert close(lf(), stats.poisson(1000).logpmf(1), 1000)
    assert close(lf(s0_rate_multiplier=5.4), stats.poisson(5400).logpmf(1), 5400)
    lf.set_data(np.zeros(0, dtype=[('x', float), ('source', int)]))
    assert lf(s0_rate_multiplier=0.) == stats.poisson(0).logpmf(0)
    lf = BinnedLogLikelihood(conf_for_test(mc=True, analysis_space=[['x', [-40, 0, 40]]], **small))
    lf.add_rate_parameter('s0')
    lf.prepare()
    lf.set_data(np.ones(100, dtype=[('x', float), ('source', int)]))
    assert almost_equal(lf(), stats.poisson(500).logpmf(100) + stats.poisson(500).logpmf(0), 1e-2)


def test_multi_bin_binned():
    """tests/test_binned_likelihood.py:41-110."""
    from blueice_b200 import BinnedLogLikelihood
    from blueice_b200.test_helpers import FixedSampleSource, almost_equal, conf_for_test, make_data
    instructions_mc = [dict(n_events=24, x=0.5, y=0.5), dict(n_events=56, x=1.5, y=0.5),
                       dict(n_events=6, x=0.5, y=2), dict(n_events=14, x=1.5, y=2)]
    data, n_mc = make_data(instructions_mc)
    conf = conf_for_test(events_per_day=42, default_source_class=FixedSampleSource, data=data,
                         analysis_space=[['x', [0, 1, 5]], ['y', [0, 1, 4]]])
    lf = BinnedLogLikelihood(conf)
    lf.add_rate_parameter('s0')
    lf.add_shape_parameter('strlen_multiplier', {1: 'x', 2: 'hi', 3: 'wha'}, base_value=1)
    lf.prepare()
    instructions_data = [dict(n_events=18, x=0.5, y=0.5), dict(n_events=70, x=1.5, y=0.5),
                         dict(n_events=4, x=0.5, y=2), dict(n_events=10, x=1.5, y=2)]
    data, _ = make_data(instructions_data)
    lf.set_data(data)
    mus = [42 / n_mc * i['n_events'] for i in instructions_mc]
    seen = [i['n_events'] for i in instructions_data]

    def expect(f):
        return np.sum([stats.poisson(f * mu).logpmf(k) for mu, k in zip(mus, seen)])

    assert almost_equal(lf(strlen_multiplier=1), expect(1))
    with pytest.raises(NotImplementedError):
        lf(compute_pdf=True, strlen_multiplier=2)
    assert almost_equal(lf(compute_pdf=False, strlen_multiplier=2), expect(2))
    assert almost_equal(lf(strlen_multiplier=2.3), expect(2.3))


def test_beeston_barlow_scenarios():
    """tests/test_BeestonBarlow.py:12-131."""
    from blueice_b200 import BinnedLogLikelihood
    from blueice_b200.likelihood import beeston_barlow_root2
    from blueice_b200.test_helpers import FixedSampleSource, almost_equal, conf_for_test, make_data
    cfg = {'model_statistical_uncertainty_handling': 'bb_single', 'bb_single_source': 0}
    # single bin
    data, _ = make_data([dict(n_events=32, x=0.5)])
    lf = BinnedLogLikelihood(conf_for_test(default_source_class=FixedSampleSource, events_per_day=32 / 5,
                                           analysis_space=[['x', [0, 1]]], data=data), likelihood_config=dict(cfg))
    lf.prepare()
    assert lf.n_model_events is not None
    lf.set_data(np.zeros(2, dtype=[('x', float), ('source', int)]))
    assert almost_equal(28.0814209, beeston_barlow_root2(np.array([32]), 0.2, np.array([1]), np.array([2])))
    assert almost_equal(lf(), stats.poisson(0.2 * (2 + 32) / (1 + 0.2)).logpmf(2))
    # four bins
    data, _ = make_data([dict(n_events=16, x=0.5), dict(n_events=30, x=1.5), dict(n_events=32, x=2.5),
                         dict(n_events=27, x=3.5)])
    lf = BinnedLogLikelihood(conf_for_test(default_source_class=FixedSampleSource, events_per_day=105 / 5,
                                           analysis_space=[['x', [0, 1, 2, 3, 4]]], data=data),
                             likelihood_config=dict(cfg))
    lf.prepare()
    obs, _ = make_data([dict(n_events=3, x=0.5), dict(n_events=5, x=1.5), dict(n_events=2, x=2.5),
                        dict(n_events=7, x=3.5)])
    lf.set_data(obs)
    dbin = np.array([3, 5, 2, 7])
    A = beeston_barlow_root2(np.array([16, 30, 32, 27]), 0.2, np.array([0.]), dbin)
    np.testing.assert_almost_equal(np.array([15.833, 29.166, 28.333, 28.333]), A, decimal=2)
    assert almost_equal(lf(), np.sum(stats.poisson(0.2 * A).logpmf(dbin)))
    # second, infinite-statistics source and a (dummy) shape parameter
    cal, _ = make_data([dict(n_events=16, x=0.5), dict(n_events=30, x=1.5), dict(n_events=32, x=2.5),
                        dict(n_events=27, x=3.5)])
    other, _ = make_data([dict(n_events=5, x=0.5), dict(n_events=7, x=1.5), dict(n_events=1, x=2.5),
                          dict(n_events=3, x=3.5)])
    conf = conf_for_test(default_source_class=FixedSampleSource, analysis_space=[['x', [0, 1, 2, 3, 4]]], dummy=1)
    conf['sources'] = [{'name': 's0', 'events_per_day': 105 / 5., 'data': cal},
                       {'name': 's1', 'events_per_day': 16., 'data': other}]
    lf = BinnedLogLikelihood(conf, likelihood_config=dict(cfg))
    lf.add_shape_parameter('dummy', (0, 1))
    lf.prepare()
    lf.set_data(obs)
    U = np.array([5, 7, 1, 3])
    A = beeston_barlow_root2(np.array([16, 30, 32, 27]), 0.2, U, dbin)
    np.testing.assert_almost_equal(np.array([14.24, 26.8070, 28.08, 26.21]), A, decimal=2)
    assert almost_equal(lf(), np.sum(stats.poisson(0.2 * A + U).logpmf(dbin)))
    # missing bb_single_source -> ValueError (likelihood.py:628-629)
    lf_bad = BinnedLogLikelihood(conf, likelihood_config={'model_statistical_uncertainty_handling': 'bb_single'})
    lf_bad.add_shape_parameter('dummy', (0, 1))
    lf_bad.prepare()
    lf_bad.set_data(obs)
    with pytest.raises(ValueError):
        lf_bad()


def test_morpher_api():
    """tests/test_morphers.py:9-35."""
    from blueice_b200 import pdf_morphers
    from blueice_b200.exceptions import NoShapeParameters
    conf = dict(hypercube_shuffle_steps=2, r_sample_points=2)
    for name, morph_class in pdf_morphers.MORPHERS.items():
        with pytest.raises(NoShapeParameters):
            morph_class(config=conf, shape_parameters=OrderedDict())
        shape_pars = OrderedDict([('bla', ({-1: -1, 0: 0, 1: 1}, None, None))])
        mr = morph_class(config=conf, shape_parameters=shape_pars)
        aps = mr.get_anchor_points(bounds=[(-1, 1)], n_models=3)
        assert isinstance(aps, list) and isinstance(aps[0], tuple)
        scalar_itp = mr.make_interpolator(lambda _: 0, extra_dims=[], anchor_models={z: None for z in aps})
        assert scalar_itp([0]) == 0
        matrix_itp = mr.make_interpolator(lambda _: 0, extra_dims=[2, 2], anchor_models={z: None for z in aps})
        np.testing.assert_array_equal(matrix_itp([0]), np.zeros((2, 2)))
        # values: bit-identical to scipy's RegularGridInterpolator
        from scipy.interpolate import RegularGridInterpolator
        rng = np.random.default_rng(0)
        table = {z: rng.random((2, 3)) for z in aps}
        itp = mr.make_interpolator(lambda m: m, extra_dims=[2, 3], anchor_models=table)
        ref = RegularGridInterpolator([np.array([-1., 0., 1.])], np.stack([table[z] for z in aps]))
        for z in (-1., -0.3, 0., 0.77, 1.):
            assert np.array_equal(itp([z]), ref([z])[0])
        with pytest.raises(ValueError):
            itp([1.5])


def test_mcsource():
    """tests/test_source.py:5-15."""
    from blueice_b200.model import Model
    from blueice_b200.test_helpers import conf_for_test
    conf = conf_for_test(mc=True, n_events_for_pdf=int(2e5))
    np.random.seed(1)
    m = Model(conf)
    s = m.sources[0]
    bins = conf['analysis_space'][0][1]
    assert s.events_per_day == 1000
    assert s.fraction_in_range > 0.9999
    assert abs(s.pdf([0]) - stats.norm.pdf(0)) < 0.02
    assert (s.pdf([bins[0]]) + s.pdf([bins[1]])) / 2 == s.pdf([(bins[0] + bins[1]) / 2])


def test_fit_scipy_and_limits():
    """tests/test_inference.py:55-93,115-127."""
    from blueice_b200 import UnbinnedLogLikelihood as LogLikelihood
    from blueice_b200.inference import bestfit_scipy, one_parameter_interval
    from blueice_b200.test_helpers import conf_for_test
    np.random.seed(3)
    lf = LogLikelihood(conf_for_test())
    lf.add_rate_parameter('s0')
    lf.set_data(lf.base_model.simulate())
    fit, ll = bestfit_scipy(lf)
    assert isinstance(fit, dict) and 's0_rate_multiplier' in fit
    assert 0.8 < fit['s0_rate_multiplier'] < 1.2
    res, ll = bestfit_scipy(lf, s0_rate_multiplier=1)
    assert len(res) == 0 and ll == lf(s0_rate_multiplier=1)
    fit_b, ll_b = bestfit_scipy(lf, batched_gradient=True)          # k+1 evaluations per gradient in one batch
    assert abs(fit_b['s0_rate_multiplier'] - fit['s0_rate_multiplier']) < 1e-3

    lf = LogLikelihood(conf_for_test())
    lf.add_rate_parameter('s0')
    lf.add_shape_parameter('some_multiplier', (0.5, 1, 1.5, 2))
    lf.prepare()
    lf.set_data(lf.base_model.simulate())
    fit, ll = bestfit_scipy(lf)
    assert 'some_multiplier' in fit and 's0_rate_multiplier' in fit
    assert lf.best_anchor()['some_multiplier'] in (0.5, 1, 1.5, 2)

    lf = LogLikelihood(conf_for_test())
    lf.add_shape_parameter('strlen_multiplier', {1: 'x', 2: 'hi', 3: 'wha'}, base_value=1)
    lf.prepare()
    lf.set_data(lf.base_model.simulate())
    fit, ll = bestfit_scipy(lf)
    assert 'strlen_multiplier' in fit

    lf = LogLikelihood(conf_for_test(n_sources=2))
    lf.add_rate_parameter('s0')
    lf.prepare()
    lf.set_data(lf.base_model.simulate())
    up = one_parameter_interval(lf, target='s0_rate_multiplier', kind='upper', bound=40)
    lo = one_parameter_interval(lf, target='s0_rate_multiplier', kind='lower', bound=0.1)
    a, b = one_parameter_interval(lf, target='s0_rate_multiplier', kind='central', bound=(0.1, 20))
    best = bestfit_scipy(lf)[0]['s0_rate_multiplier']
    assert lo < best < up and a < best < b


def test_likelihood_sum_composes_batches():
    from blueice_b200 import LogLikelihoodSum, UnbinnedLogLikelihood
    from blueice_b200.test_helpers import conf_for_test
    lfs = []
    for seed in (1, 2):
        np.random.seed(seed)
        lf = UnbinnedLogLikelihood(conf_for_test())
        lf.add_rate_parameter('s0')
        lf.add_shape_parameter('some_multiplier', (0.5, 1, 2))
        lf.prepare()
        lf.set_data(lf.base_model.simulate())
        lfs.append(lf)
    total = LogLikelihoodSum(lfs, likelihood_weights=[1, 0.5])
    pts = np.array([[1.0, 1.0], [1.2, 0.7], [0.9, 1.9]])
    names = ['s0_rate_multiplier', 'some_multiplier']
    got = total.batch(pts, names)
    for row, g in zip(pts, got):
        kw = dict(zip(names, [float(v) for v in row]))
        assert g == total(**kw)
        assert g == lfs[0](**kw) + 0.5 * lfs[1](**kw)


def test_profile_scan_lockstep_matches_per_hypothesis_fits():
    """inference.profile_scan: conditional fits of a grid of hypotheses in lock step (two ll.batch passes per
    iteration) against bestfit_scipy with the hypothesis fixed (the loop of one_parameter_interval)."""
    from blueice_b200.inference import bestfit_scipy, profile_scan
    ll, d, names = wl.c2_api(n_sources=2, n_shape=2, anchors=(-2., -1., 0., 1., 2.), bins=(40, 30), n_events=3000, seed=3)
    values = np.linspace(0.6, 1.4, 9)
    prof, cond = profile_scan(ll, 'sig_rate_multiplier', values)
    assert prof.shape == (9,) and set(cond.keys()) == {'bg_rate_multiplier', 'shift1', 'shift2'}
    for h in (0, 4, 8):
        _, ref = bestfit_scipy(ll, pass_bounds_to_minimizer=True, minimize_kwargs=dict(method='L-BFGS-B'),
                               sig_rate_multiplier=float(values[h]))
        assert abs(prof[h] - ref) <= 5e-2, (h, prof[h], ref)          # kinks at the anchors: neighbouring local optima
        kw = {n: float(cond[n][h]) for n in cond}
        assert ll(sig_rate_multiplier=float(values[h]), **kw) == prof[h]
    # 3000 events against ~1e5 expected: the profile falls monotonically with the signal rate
    assert np.all(np.diff(prof) < 0)


# ------------------------------------------------------------------------------------------------
# LogLikelihoodReParam (blueice/likelihood.py:715-864; the reference's tests/test_likelihood_reparam.py)
# ------------------------------------------------------------------------------------------------
def _reparam_pair():
    from copy import deepcopy
    from blueice_b200.likelihood import LogLikelihoodReParam, UnbinnedLogLikelihood
    from blueice_b200.test_helpers import BASE_CONV_CONFIG, conf_for_reparam_test
    lf_old = UnbinnedLogLikelihood(conf_for_reparam_test(events_per_day=1.))
    for name in ("op0", "op1", "op2"):
        lf_old.add_rate_parameter(name)
    lf_old.prepare()
    return lf_old, LogLikelihoodReParam(lf_old, deepcopy(BASE_CONV_CONFIG))


def test_golden_reparam():
    """Values of the unmodified reference's LogLikelihoodReParam (tests/golden/make_golden.py reparam)."""
    g = load_golden('reparam')
    _, lf = _reparam_pair()
    d = np.zeros(len(g['x']), dtype=[('x', float), ('source', int)])
    d['x'] = g['x']
    lf.set_data(d)
    n = len(d)
    got = np.array([lf(np0=float(a), np1=float(b)) for a, b in g['points']])
    assert np.all(np.abs(got - g['logl']) <= 1e-9 * n)
    assert np.all(np.abs(got - g['logl']) <= 1e-13 * (np.abs(g['logl']) + n))
    assert np.array_equal(lf.batch(g['points'], ['np0', 'np1']), got)               # batch == scalar, bit for bit
    only = lf.batch(g['points'][:, :1], ['np0'])
    assert np.all(np.abs(only - g['logl_only_np0']) <= 1e-9 * n)
    assert abs(lf() - g['default'][0]) <= 1e-9 * n


def test_reparam_likelihood_value():
    """tests/test_likelihood_reparam.py:9-41."""
    _, lf = _reparam_pair()
    d = np.zeros(3, dtype=[('x', float), ('source', int)])
    lf.set_data(d)
    for v in (1, 2, 3):
        total = v ** 2 + v ** 2 + v * v
        expect = -total + 3 * np.log(total) + 3 * stats.norm.logpdf(0)
        assert np.isclose(lf(np0=v, np1=v), expect, atol=1e-08)


def test_reparam_likelihoods_before_after():
    """tests/test_likelihood_reparam.py:44-72: the wrapped and the re-parameterised likelihood agree; here exactly."""
    lf_old, lf = _reparam_pair()
    np.random.seed(2)
    d = lf.base_model.simulate()
    lf.set_data(d)
    lf_old.set_data(d)
    assert lf() == lf_old()
    assert lf(np0=2) == lf_old(op0_rate_multiplier=4, op2_rate_multiplier=2)
    assert lf(np1=2) == lf_old(op1_rate_multiplier=4, op2_rate_multiplier=2)
    assert lf(np0=2, np1=2) == lf_old(op0_rate_multiplier=4, op1_rate_multiplier=4, op2_rate_multiplier=4)
    d2 = lf.base_model.simulate(dict(np0=2.0), livetime_days=3.0)                   # _simulate through the converter
    assert len(d2) >= 0
    best, ll_max = lf.bestfit_scipy()
    assert set(best) == {'np0', 'np1'} and np.isfinite(ll_max)


def test_lean_paths_equal_the_general_path(monkeypatch):
    """ll(**kw) and ll.batch through the cached plans (columns written straight into the pinned staging buffer, kernels
    reading / writing pinned memory) against the general vectorised path (BI_SCALAR_FAST=0, BI_DIRECT_IO=0): same bits,
    with priors, live-time scaling, missing columns, out-of-range and unphysical rows."""
    import importlib
    from blueice_b200 import engine as engine_mod
    results = {}
    for fast in ("1", "0"):
        monkeypatch.setenv("BI_SCALAR_FAST", fast)
        monkeypatch.setattr(engine_mod, "_DIRECT_IO", fast == "1")
        ll, d, names = wl.c2_api(2, 2, (-1., 0., 1.), (40, 30), n_events=20000, seed=11)
        ll.rate_parameters['bg'] = sps.norm(1, 0.2).logpdf
        ll.shape_parameters['shift1'] = (ll.shape_parameters['shift1'][0], sps.norm(0, 1).logpdf, None)
        ll.set_data(d)
        zs, mult = wl.scan_points(700, 2, 2, seed=12, z_range=(-1., 1.))
        zs[5, 0] = 3.0
        mult[6, 1] = -1.0
        table = np.column_stack([mult, zs])
        out = [ll.batch(table, names) for _ in range(3)]
        assert np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[2])
        part = ll.batch(table[:, [0, 2]], [names[0], names[2]], livetime_days=2.5)
        one = [ll(**dict(zip(names, [float(v) for v in table[i]]))) for i in (0, 5, 6, 9)]
        lt = ll(livetime_days=0.5, **{names[0]: 1.3})
        results[fast] = (out[0], part, np.array(one), lt, ll.batch(table[:1], names), ll.batch(table[:64], names))
    for a, b in zip(results["1"], results["0"]):
        assert np.array_equal(a, b)
    assert results["1"][0][5] == -np.inf and results["1"][0][6] == -np.inf
    assert np.array_equal(results["1"][2], results["1"][0][[0, 5, 6, 9]])


def test_long_contraction_model_through_the_api():
    """3 shape parameters (8 corners per hypercube cell) x 5 sources = 40 contraction terms: ll.batch runs the K-chunk form of
    K2 (more than 32 terms).  Scalar calls == batch rows bit for bit (the scalar path is the same kernel at P = 1), and the
    batch agrees with the oracle evaluated on the engine's own per-event anchor tensor (likelihood.py:355-356,678-690)."""
    from blueice_b200 import UnbinnedLogLikelihood
    from blueice_b200.test_helpers import conf_for_test
    from oracle.pipeline import UnbinnedOracle
    lf = UnbinnedLogLikelihood(conf_for_test(n_sources=5, events_per_day=300.))
    lf.add_shape_parameter('mu', (-0.5, 0., 0.5))
    lf.add_shape_parameter('sigma', (0.8, 1., 1.3))
    lf.add_shape_parameter('some_multiplier', (0.5, 1., 2.))
    for i in range(5):
        lf.add_rate_parameter('s%d' % i)
    lf.prepare()
    d = lf.base_model.simulate()
    lf.set_data(d)
    n = len(d)
    eng = lf._engine
    assert eng.n_terms == 40 and eng.uses_mma()
    names = lf.parameter_names()
    rng = np.random.default_rng(3)
    P = 300
    mult = rng.uniform(0.5, 2, (P, 5))
    zs = np.column_stack([rng.uniform(-0.5, 0.5, P), rng.uniform(0.8, 1.3, P), rng.uniform(0.5, 2., P)])
    zs[:3] = [[-0.5, 0.8, 0.5], [0., 1., 1.], [0.5, 1.3, 2.]]              # on anchors
    table = np.column_stack([mult, zs])
    got = lf.batch(table, names)
    assert np.all(np.isfinite(got))
    for i in (0, 1, 2, 17, 299):
        assert lf(**dict(zip(names, [float(v) for v in table[i]]))) == got[i]
    assert lf() == lf.batch(np.array([[1.] * 5 + [0., 1., 1.]]), names)[0]
    axes = [np.asarray(a, dtype=float) for a in eng.grid.axes]
    shape = [len(a) for a in axes]
    ps = eng.ps_anchor[:, :, :n].cpu().numpy().reshape(shape + [5, n])
    orc = UnbinnedOracle(axes, eng.mus_anchor_host.reshape(shape + [5])).set_ps(ps)
    assert_vector_close(got, orc.batch(zs, mult), n)
    table[7, 5] = 0.6                                                   # mu beyond its last anchor -> -inf
    table[8, 0] = -1.0                                                  # unphysical rate -> -inf
    out = lf.batch(table, names)
    assert out[7] == -np.inf and out[8] == -np.inf
    keep = np.ones(P, dtype=bool)
    keep[[7, 8]] = False
    assert np.array_equal(out[keep], got[keep])
