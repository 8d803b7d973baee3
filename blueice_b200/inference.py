"""Fitting and interval helpers that drive the likelihood (callers of the hot path).

Same public surface as blueice/inference.py (`make_objective` :57-124, `bestfit_scipy` :131-178,
`one_parameter_interval` :332-389, `best_anchor` :34-54, `plot_likelihood_ratio` :392-443); every one
of them is also bound as a method of the likelihood classes.  They stay Python: each evaluation is
one `lf(**kwargs)` -> one device pass.  Additions that exploit the batched entry point:

  * `make_objective(...)` returns an objective with a `.batch(array[P, k])` attribute (one device
    pass for P points);
  * `bestfit_scipy(..., batched_gradient=True)` hands scipy a 2-point finite-difference `jac`
    evaluated as ONE batch of k+1 points per step (the k+1 evaluations BFGS would otherwise issue
    one at a time); default False keeps the reference's call sequence unchanged.

bestfit_minuit / bestfit_emcee (optional iminuit 1.x / emcee dependencies)
are thin wrappers kept for API completeness; they only call the objective.
"""
import warnings
from collections import OrderedDict
from copy import deepcopy

import numpy as np
from scipy import stats
from scipy.optimize import brentq, minimize

from .exceptions import NoOpimizationNecessary, OptimizationFailed

DEFAULT_BESTFIT_ROUTINE = 'scipy'
_RATE_SUFFIX = '_rate_multiplier'

__all__ = ['best_anchor', 'make_objective', 'bestfit_scipy', 'bestfit_minuit', 'plot_likelihood_ratio',
           'one_parameter_interval', 'bestfit_emcee']


def best_anchor(lf):
    """Shape-parameter dict of the anchor model with the highest likelihood (one batched pass)."""
    if not len(lf.shape_parameters):
        return dict()
    names = list(lf.shape_parameters.keys())
    anchors = list(lf.anchor_models.keys())
    values = lf.batch(np.asarray(anchors, dtype=np.float64), names)
    return dict(zip(names, anchors[int(np.argmax(values))]))


def make_objective(lf, guess=None, minus=True, rates_in_log_space=False, **kwargs):
    """Positional-argument objective for optimisers.

    :param kwargs: fixed parameter values (not fitted)
    :param guess: {name: starting value} for floating parameters (default: base value)
    :param minus: multiply the log likelihood by -1 (minimisers want this)
    :param rates_in_log_space: let the optimiser see log10 of the rate multipliers
    :returns: (f, names, guesses, bounds); f takes one array of the floating parameters in `names` order
    """
    guess = guess or {}
    names, guesses, bounds = [], [], []
    for source_name in lf.rate_parameters.keys():
        key = source_name + _RATE_SUFFIX
        if key in kwargs:
            continue
        start = guess.get(key, 1)
        names.append(key)
        guesses.append(np.log10(start) if rates_in_log_space else start)
        bounds.append((None, None) if rates_in_log_space else (0, None))
    for setting, (_, _, base_value) in lf.shape_parameters.items():
        if setting in kwargs:
            continue
        start = guess.get(setting)
        if start is None:
            start = lf.pdf_base_config.get(setting)
            if not isinstance(start, (int, float)):
                start = base_value
        names.append(setting)
        guesses.append(start)
        bounds.append(lf.get_bounds(setting))
    if not names:
        raise NoOpimizationNecessary("There are no parameters to fit, no optimization is necessary")

    sign = -1 if minus else 1
    is_log_rate = [rates_in_log_space and n.endswith(_RATE_SUFFIX) for n in names]

    def objective(args):
        call = {n: (10 ** args[i] if is_log_rate[i] else args[i]) for i, n in enumerate(names)}
        call.update(kwargs)
        return lf(**call) * sign

    def objective_batch(points):
        points = np.array(points, dtype=np.float64).reshape(-1, len(names))
        for i, flag in enumerate(is_log_rate):
            if flag:
                points[:, i] = 10 ** points[:, i]
        fixed = list(kwargs.keys())
        cols = np.empty((len(points), len(names) + len(fixed)))
        cols[:, :len(names)] = points
        for j, key in enumerate(fixed):
            cols[:, len(names) + j] = kwargs[key]
        return lf.batch(cols, names + fixed) * sign

    if hasattr(lf, 'batch'):
        objective.batch = objective_batch
    return objective, names, np.array(guesses), bounds


def _forward_difference_jac(f, step=1.4901161193847656e-08):
    """2-point forward differences with scipy's default step rule, k+1 points in one batch."""
    def jac(x):
        x = np.asarray(x, dtype=np.float64)
        h = step * np.where(x >= 0, 1.0, -1.0) * np.maximum(1.0, np.abs(x))
        pts = np.vstack([x] + [x + h[i] * np.eye(len(x))[i] for i in range(len(x))])
        vals = f.batch(pts)
        return (vals[1:] - vals[0]) / ((x + h) - x)
    return jac


def bestfit_scipy(lf, minimize_kwargs=None, rates_in_log_space=False, pass_bounds_to_minimizer=False,
                  batched_gradient=False, **kwargs):
    """Maximise lf over the parameters not fixed in kwargs with scipy.optimize.minimize.

    Returns ({parameter: best fit}, maximum log likelihood).  Falls back to Nelder-Mead once if the
    first attempt reports failure; raises OptimizationFailed if that fails too.
    """
    minimize_kwargs = {} if minimize_kwargs is None else minimize_kwargs
    try:
        f, names, guess, bounds = lf.make_objective(minus=True, rates_in_log_space=rates_in_log_space, **kwargs)
    except NoOpimizationNecessary:
        return {}, lf(**kwargs)

    use_bounds = bounds if pass_bounds_to_minimizer else None
    first_kwargs = dict(minimize_kwargs)
    if batched_gradient and hasattr(f, 'batch') and 'jac' not in first_kwargs:
        first_kwargs['jac'] = _forward_difference_jac(f)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        res = minimize(f, guess, bounds=use_bounds, **first_kwargs)
        if not res.success:
            retry_kwargs = deepcopy(minimize_kwargs)
            retry_kwargs.pop('method', None)
            res = minimize(f, guess, bounds=use_bounds, method='Nelder-Mead', **retry_kwargs)
            if not res.success:
                raise OptimizationFailed("Optimization failure: ", res)

    best = res.x if len(names) != 1 else [res.x.item()]
    fit = OrderedDict()
    for i, name in enumerate(names):
        fit[name] = 10 ** best[i] if (rates_in_log_space and name.endswith(_RATE_SUFFIX)) else best[i]
    return fit, -res.fun


def bestfit_minuit(lf, minimize_kwargs=None, rates_in_log_space=False, **kwargs):
    """Minimise with iminuit 1.x (optional dependency)."""
    from iminuit import Minuit
    from iminuit.util import make_func_code
    minimize_kwargs = {} if minimize_kwargs is None else minimize_kwargs
    minimize_kwargs.setdefault('print_level', 0)
    minimize_kwargs.setdefault('pedantic', False)
    try:
        f, names, guess, bounds = lf.make_objective(minus=True, rates_in_log_space=rates_in_log_space, **kwargs)
    except NoOpimizationNecessary:
        return {}, lf(**kwargs)
    setup = minimize_kwargs
    for i, name in enumerate(names):
        setup[name] = guess[i]
        setup['limit_' + name] = bounds[i]
    setup['errordef'] = 0.5

    class _Wrapped:
        def __init__(self, func, arg_names):
            self.func = func
            self.func_code = make_func_code(arg_names)

        def __call__(self, *args):
            return self.func(args)

    m = Minuit(_Wrapped(f, names), **setup)
    m.migrad()
    fit = {k: v for k, v in m.values.items()}
    for k, v in m.errors.items():
        fit[k + '_error'] = v
    return fit, -1 * m.fval


def bestfit_emcee(ll, quiet=False, return_errors=False, return_samples=False, n_walkers=40, n_steps=200,
                  n_burn_in=100, n_threads=1, **kwargs):
    """Median of an emcee ensemble as point estimate (optional dependency)."""
    import emcee
    f, names, guess, _ = ll.make_objective(minus=False, **kwargs)
    n_dim = len(guess)
    start = np.array([np.random.uniform(0.95, 1.05, size=n_dim) * guess for _ in range(n_walkers)])
    sampler = emcee.EnsembleSampler(n_walkers, n_dim, f, threads=n_threads)
    sampler.run_mcmc(start, n_steps)
    samples = sampler.chain[:, n_burn_in:, :].reshape((-1, n_dim))
    if not quiet:
        print("Mean acceptance fraction: {0:.3f}".format(np.mean(sampler.acceptance_fraction)))
    centre = np.median(samples, axis=0)
    fit = OrderedDict(zip(names, centre))
    best_ll = ll(**fit)
    if return_errors:
        lo, hi = np.percentile(samples, 100 * stats.norm.cdf([-1, 1]), axis=0)
        return fit, best_ll, OrderedDict(zip(names, (hi - lo) / 2))
    if return_samples:
        return fit, best_ll, samples
    return fit, best_ll


def _get_bestfit_routine(key):
    if callable(key):
        return key
    return BESTFIT_ROUTINES[DEFAULT_BESTFIT_ROUTINE if key is None else key]


def one_parameter_interval(lf, target, bound, confidence_level=0.9, kind='upper', bestfit_routine=None,
                           t_ppf=None, **kwargs):
    """Profile-likelihood interval (upper / lower / central) on parameter `target`.

    Assumes the likelihood ratio is chi2(1) distributed (Wilks) unless t_ppf(hypothesis, level) is given.
    bound: line-search bound (2-tuple for kind='central'); kwargs go to the fitting routine.
    """
    fit = _get_bestfit_routine(bestfit_routine)
    if target is None:
        target = lf.source_list[-1] + _RATE_SUFFIX
    best, max_ll = fit(lf, **kwargs)
    global_best = best[target]

    def t(hypothesis, critical_quantile):
        if t_ppf is None:
            # norm.ppf(CL)**2 == chi2(1).ppf(2*CL - 1): one-sided vs two-sided quoting of Wilks' theorem
            critical = stats.norm.ppf(critical_quantile) ** 2
        else:
            critical = t_ppf(hypothesis, critical_quantile)
        one_sided_trivial = (kind == 'upper' and hypothesis <= global_best) or \
                            (kind == 'lower' and hypothesis >= global_best)
        if one_sided_trivial:
            statistic = 0
        else:
            conditional = {target: hypothesis}
            conditional.update(kwargs)
            _, ll = fit(lf, **conditional)
            statistic = 2 * (max_ll - ll)
        return statistic - critical

    if kind == 'central':
        lo = brentq(t, bound[0], global_best, args=((1 - confidence_level) / 2,))
        hi = brentq(t, global_best, bound[1], args=(1 - (1 - confidence_level) / 2,))
        return lo, hi
    if kind == 'lower':
        return brentq(t, bound, global_best, args=(1 - confidence_level,))
    if kind == 'upper':
        return brentq(t, global_best, bound, args=(confidence_level,))
    raise ValueError("kind must be 'upper', 'lower' or 'central'")


def plot_likelihood_ratio(lf, *space, vmax=15, bestfit_routine=None, plot_kwargs=None, **kwargs):
    """Plot the profile -log likelihood ratio over 1 or 2 parameters (needs matplotlib)."""
    import matplotlib.pyplot as plt
    fit = _get_bestfit_routine(bestfit_routine)
    plot_kwargs = plot_kwargs or {}
    label = "-Log likelihood ratio"
    if len(space) == 1:
        dim, xs = space[0]
        ll = np.array([fit(lf, **dict(kwargs, **{dim: x}))[1] for x in xs])
        plt.plot(xs, ll.max() - ll, **plot_kwargs)
        plt.ylim(0, vmax)
        plt.ylabel(label)
        plt.xlabel(dim)
        plt.xlim(xs.min(), xs.max())
    elif len(space) == 2:
        (d0, xs), (d1, ys) = space
        ll = np.array([[fit(lf, **dict(kwargs, **{d0: x, d1: y}))[1] for y in ys] for x in xs])
        grid_x, grid_y = np.meshgrid(xs, ys)
        plt.pcolormesh(grid_x, grid_y, (np.nanmax(ll) - ll).T, vmax=vmax, **plot_kwargs)
        plt.colorbar(label=label)
        plt.xlabel(d0)
        plt.ylabel(d1)
    else:
        raise ValueError("Can't handle %d dimensions" % len(space))


BESTFIT_ROUTINES = dict(scipy=bestfit_scipy, minuit=bestfit_minuit, emcee=bestfit_emcee)
