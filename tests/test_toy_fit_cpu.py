"""Host logic of the lock-step toy fitter (inference.bestfit_toys) on the CPU: the per-toy likelihood is played by
the oracle (tests may use it), so this checks the optimiser -- projected BFGS with batched central differences and
batched trial steps -- against scipy's L-BFGS-B toy by toy, the way the reference fits toys."""
from collections import OrderedDict

import numpy as np
from scipy.optimize import minimize

import bench_workloads as wl
from blueice_b200 import inference
from oracle import toys as otoys
from oracle.pipeline import UnbinnedOracle


class OracleToyLikelihood(object):
    """The attributes and the batch_toys(toy_index=...) call bestfit_toys uses, backed by per-toy oracles."""

    def __init__(self, n_toys=5, seed=8, livetime=0.02):
        anchors = (-2., -1., 0., 1., 2.)
        axes, edges, templates, mus = wl.c2_arrays(2, 2, anchors, (40, 30))
        vol = np.outer(np.diff(edges[0]), np.diff(edges[1]))
        cdf = np.vstack([np.cumsum((templates[2, 2, s] * vol).ravel()) for s in range(2)])
        cdf /= cdf[:, -1:]
        counts = np.random.default_rng(1).poisson(mus[2, 2] * livetime, size=(n_toys, 2))
        coords, _, offsets = otoys.toy_events(edges, cdf, counts, seed=seed)
        self.oracles = [UnbinnedOracle(axes, mus * livetime).set_data_from_templates(
            templates, edges, [c[offsets[t]:offsets[t + 1]] for c in coords]) for t in range(n_toys)]
        self.n_toys = n_toys
        self.rate_parameters = OrderedDict([('bg', None), ('sig', None)])
        self.shape_parameters = OrderedDict([('shift1', (None, None, 0.)), ('shift2', (None, None, 0.))])
        self.pdf_base_config = {'shift1': 0., 'shift2': 0.}
        self.calls = 0

    def get_bounds(self, name):
        return (-2., 2.)

    def batch_toys(self, cols, names, livetime_days=None, toy_index=None):
        cols = np.asarray(cols, dtype=float)
        toy_index = np.arange(self.n_toys) if toy_index is None else toy_index
        im = [names.index('bg_rate_multiplier'), names.index('sig_rate_multiplier')]
        iz = [names.index('shift1'), names.index('shift2')]
        self.calls += 1
        return np.array([self.oracles[t](row[iz], row[im]) for t, row in zip(toy_index, cols)])


def test_bestfit_toys_reaches_the_scipy_optimum_of_every_toy():
    ll = OracleToyLikelihood()
    fit, maxll, info = inference.bestfit_toys(ll)
    assert info['converged'].all() and list(fit.keys()) == ['bg_rate_multiplier', 'sig_rate_multiplier', 'shift1', 'shift2']
    # lock step: a handful of batched passes per iteration, not one call per toy and point
    assert ll.calls <= 2 * info['iterations'] + 2
    for t in range(ll.n_toys):
        res = minimize(lambda x: -ll.oracles[t](x[2:], x[:2]), [1, 1, 0, 0], method='L-BFGS-B',
                       bounds=[(0, None), (0, None), (-2, 2), (-2, 2)])
        assert maxll[t] >= -res.fun - 1e-4, (t, maxll[t], -res.fun)
        assert abs(maxll[t] + res.fun) <= 1e-2
        assert maxll[t] == ll.oracles[t]([fit['shift1'][t], fit['shift2'][t]],
                                         [fit['bg_rate_multiplier'][t], fit['sig_rate_multiplier'][t]])


def test_bestfit_toys_fixed_parameters_and_bounds():
    ll = OracleToyLikelihood(n_toys=3)
    free, free_ll, _ = inference.bestfit_toys(ll)
    cond, cond_ll, _ = inference.bestfit_toys(ll, sig_rate_multiplier=0.5, shift2=0.0)
    assert list(cond.keys()) == ['bg_rate_multiplier', 'shift1']
    assert np.all(cond_ll <= free_ll + 1e-6)                           # a conditional maximum cannot beat the global one
    assert np.all((cond['shift1'] >= -2) & (cond['shift1'] <= 2) & (cond['bg_rate_multiplier'] >= 0))
    # a guess outside the box is clipped into it, not rejected
    g, gll, _ = inference.bestfit_toys(ll, guess={'shift1': 5.0})
    assert np.all(np.abs(gll - free_ll) <= 1e-2)


def test_profile_scan_matches_conditional_scipy_fits():
    """All hypotheses of a profile-likelihood scan fitted in lock step == one scipy fit per hypothesis."""
    toy = OracleToyLikelihood(n_toys=1)
    orc = toy.oracles[0]

    class OneDataset(object):
        rate_parameters, shape_parameters, pdf_base_config = toy.rate_parameters, toy.shape_parameters, toy.pdf_base_config
        get_bounds = staticmethod(lambda name: (-2., 2.))

        def batch(self, cols, names, livetime_days=None):
            cols = np.asarray(cols, dtype=float)
            im = [names.index('bg_rate_multiplier'), names.index('sig_rate_multiplier')]
            iz = [names.index('shift1'), names.index('shift2')]
            return np.array([orc(row[iz], row[im]) for row in cols])

    values = np.array([0.0, 0.5, 1.0, 1.6])
    prof, cond = inference.profile_scan(OneDataset(), 'sig_rate_multiplier', values)
    assert list(cond.keys()) == ['bg_rate_multiplier', 'shift1', 'shift2'] and prof.shape == (4,)
    for h, v in enumerate(values):
        bounds = [(0, None), (-2, 2), (-2, 2)]
        res = minimize(lambda x: -orc(x[1:], [x[0], v]), [1, 0, 0], method='L-BFGS-B', bounds=bounds)
        # the multilinear morph has kinks at the anchors: two optimisers may settle in neighbouring cells, a few 1e-3
        # apart in log likelihood; what must hold is (a) closeness and (b) that scipy cannot improve on our point
        assert abs(prof[h] + res.fun) <= 2e-2, (h, prof[h], -res.fun)
        polish = minimize(lambda x: -orc(x[1:], [x[0], v]), [cond[n][h] for n in cond], method='L-BFGS-B', bounds=bounds)
        assert -polish.fun <= prof[h] + 1e-6, (h, prof[h], -polish.fun)
    # everything else fixed: the scan is one batch evaluation
    flat, none = inference.profile_scan(OneDataset(), 'sig_rate_multiplier', values, bg_rate_multiplier=1.0,
                                        shift1=0.0, shift2=0.0)
    assert len(none) == 0 and flat[2] == orc([0., 0.], [1., 1.])


def test_interval_scan_matches_the_sequential_interval():
    """one_parameter_interval_scan (grids of conditional fits in lock step) lands where brentq over sequential scipy
    fits does (the reference's one_parameter_interval, inference.py:332-389)."""
    from scipy import stats
    from scipy.optimize import brentq
    toy = OracleToyLikelihood(n_toys=1)
    orc = toy.oracles[0]

    class OneDataset(object):
        rate_parameters, shape_parameters, pdf_base_config = toy.rate_parameters, toy.shape_parameters, toy.pdf_base_config
        get_bounds = staticmethod(lambda name: (-2., 2.))
        source_list = ['bg', 'sig']

        def parameter_names(self):
            return ['bg_rate_multiplier', 'sig_rate_multiplier', 'shift1', 'shift2']

        def batch(self, cols, names, livetime_days=None):
            cols = np.asarray(cols, dtype=float)
            im = [names.index('bg_rate_multiplier'), names.index('sig_rate_multiplier')]
            iz = [names.index('shift1'), names.index('shift2')]
            return np.array([orc(row[iz], row[im]) for row in cols])

    lf = OneDataset()
    # shift parameters fixed: a smooth 2-parameter problem, so both methods must agree closely
    limit = inference.one_parameter_interval_scan(lf, 'sig_rate_multiplier', 3.0, confidence_level=0.9, kind='upper',
                                                  shift1=0.0, shift2=0.0)
    bounds = [(0, None)]
    free = minimize(lambda x: -orc([0., 0.], [x[0], x[1]]), [1, 1], method='L-BFGS-B', bounds=[(0, None), (0, None)])

    def t(h):
        cond = minimize(lambda x: -orc([0., 0.], [x[0], h]), [1], method='L-BFGS-B', bounds=bounds)
        return 2 * (-free.fun + cond.fun) - stats.norm.ppf(0.9) ** 2

    ref = brentq(t, free.x[1], 3.0)
    assert abs(limit - ref) <= 2e-3 * ref, (limit, ref)
    lo, hi = inference.one_parameter_interval_scan(lf, 'sig_rate_multiplier', (0.0, 3.0), confidence_level=0.68,
                                                   kind='central', shift1=0.0, shift2=0.0)
    assert lo < free.x[1] < hi
