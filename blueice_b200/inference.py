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
           'one_parameter_interval', 'bestfit_emcee',
           # batched drivers (not in the reference): many fits advanced in lock step on the device
           'bestfit_toys', 'profile_scan', 'one_parameter_interval_scan']


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


def _lockstep_bfgs(minus_ll, x0, lo, hi, max_iter=60, gtol=1e-3, ftol=1e-10, fd_step=1e-5,
                   line_search=(2.0, 1.0, 0.5, 0.25, 0.1, 0.02)):
    """F independent box-constrained minimisations advanced in lock step.

    minus_ll(points [Q, k], fit_index [Q]) -> values [Q] evaluates point q of fit fit_index[q]; every iteration makes
    two calls for all unconverged fits together: the 2k central-difference points (which also give the diagonal
    curvature that scales the first step and every restart) and the trial steps of the line search.  Projected BFGS:
    parameters pinned at a bound by the gradient do not move.  Returns (x [F, k], f [F], converged [F], iterations)."""
    x = np.clip(np.array(x0, dtype=np.float64).reshape(-1, len(lo)), lo, hi)
    T, k = x.shape
    f = minus_ll(x, np.arange(T))
    if not np.all(np.isfinite(f)):
        raise OptimizationFailed("the starting point has a non-finite likelihood for %d of the %d fits"
                                 % (int((~np.isfinite(f)).sum()), T))
    H = np.broadcast_to(np.eye(k), (T, k, k)).copy()
    fresh = np.ones(T, dtype=bool)                 # H must be (re)built from the diagonal second differences
    g_old, x_old = np.zeros((T, k)), x.copy()
    converged = np.zeros(T, dtype=bool)
    stalled = np.zeros(T, dtype=int)
    small_gains = np.zeros(T, dtype=int)
    eye = np.eye(k)
    steps = np.asarray(line_search, dtype=np.float64)
    it = 0
    for it in range(1, max_iter + 1):
        act = np.flatnonzero(~converged)
        if not len(act):
            break
        A = len(act)
        xa, fa = x[act], f[act]
        # ---- central differences inside the box (one-sided where a bound is closer than the step)
        h = fd_step * np.maximum(1.0, np.abs(xa))
        up = np.minimum(xa + h, hi)
        dn = np.maximum(xa - h, lo)
        pts = np.repeat(xa[:, None, :], 2 * k, axis=1)                 # [A, 2k, k]
        for j in range(k):
            pts[:, 2 * j, j] = up[:, j]
            pts[:, 2 * j + 1, j] = dn[:, j]
        vals = minus_ll(pts.reshape(A * 2 * k, k), np.repeat(act, 2 * k)).reshape(A, 2 * k)
        f_up, f_dn = vals[:, 0::2], vals[:, 1::2]
        g = (f_up - f_dn) / (up - dn)
        g = np.where(np.isfinite(g), g, 0.0)
        # diagonal curvature from the same points (used to scale the first step and every restart)
        with np.errstate(invalid='ignore', divide='ignore'):
            curv = 2.0 * ((f_up - fa[:, None]) / (up - xa) - (fa[:, None] - f_dn) / (xa - dn)) / (up - dn)
        curv = np.where(np.isfinite(curv) & (curv > 1e-8), curv, np.nan)
        # ---- BFGS update of the inverse Hessian with the step taken last iteration
        sv = xa - x_old[act]
        yv = g - g_old[act]
        sy = np.einsum('ij,ij->i', sv, yv)
        ok = (~fresh[act]) & (sy > 1e-10 * np.sqrt(np.einsum('ij,ij->i', yv, yv) * np.einsum('ij,ij->i', sv, sv)))
        rho = np.where(ok, 1.0 / np.where(ok, sy, 1.0), 0.0)
        V = eye[None] - rho[:, None, None] * sv[:, :, None] * yv[:, None, :]
        Hn = np.matmul(np.matmul(V, H[act]), np.transpose(V, (0, 2, 1))) + rho[:, None, None] * sv[:, :, None] * sv[:, None, :]
        Ha = np.where(ok[:, None, None], Hn, H[act])
        scale = np.where(np.isnan(curv), np.nanmedian(np.where(np.isnan(curv), np.nan, 1.0 / curv), axis=1, keepdims=True),
                         1.0 / curv)
        scale = np.where(np.isfinite(scale), scale, 1.0)
        H0 = scale[:, :, None] * eye[None]
        Ha = np.where(fresh[act][:, None, None], H0, Ha)
        # ---- projected direction: parameters pinned at a bound by the gradient do not move
        pinned = ((xa <= lo) & (g > 0)) | ((xa >= hi) & (g < 0))
        gp = np.where(pinned, 0.0, g)
        d = np.where(pinned, 0.0, -np.matmul(Ha, gp[:, :, None])[:, :, 0])
        uphill = np.einsum('ij,ij->i', d, gp) >= 0
        d = np.where(uphill[:, None], np.where(pinned, 0.0, -scale * gp), d)
        Ha = np.where(uphill[:, None, None], H0, Ha)
        H[act] = Ha
        # ---- trial steps, all in one batch
        n_ls = len(steps)
        trial = np.clip(xa[:, None, :] + steps[None, :, None] * d[:, None, :], lo, hi)
        ft = minus_ll(trial.reshape(A * n_ls, k), np.repeat(act, n_ls)).reshape(A, n_ls)
        ft = np.where(np.isfinite(ft), ft, np.inf)
        best = np.argmin(ft, axis=1)
        fb = ft[np.arange(A), best]
        better = fb < fa
        g_old[act] = g
        x_old[act] = xa
        gain = np.where(better, fa - fb, 0.0)
        x[act[better]] = trial[np.arange(A), best][better]
        f[act[better]] = fb[better]
        was_fresh = fresh[act]
        small = better & (gain <= ftol * np.maximum(1.0, np.abs(fa)))
        # a failed or nearly useless line search restarts from the scaled diagonal; only when such a restart does
        # not help either does it count towards convergence
        fresh[act] = (~better) | (small & ~was_fresh)
        stalled[act] = np.where(better, 0, stalled[act] + np.where(was_fresh, 1, 0))
        small_gains[act] = np.where(small & was_fresh, small_gains[act] + 1, np.where(small, small_gains[act], 0))
        small_gradient = np.max(np.abs(gp * np.sqrt(scale)), axis=1) <= gtol     # gradient in units of the curvature
        done = small_gradient | (small_gains[act] >= 2) | (stalled[act] >= 2)
        converged[act[done]] = True
    return x, f, converged, it


def bestfit_toys(lf, guess=None, livetime_days=None, max_iter=60, gtol=1e-3, ftol=1e-10, fd_step=1e-5,
                 line_search=(2.0, 1.0, 0.5, 0.25, 0.1, 0.02), **kwargs):
    """Maximise the likelihood of EVERY toy of lf.set_toy_data over the parameters not fixed in kwargs, all toys
    in lock step on the device (the per-toy fits of a Neyman construction; not in the reference, whose callers run
    bestfit_scipy toy by toy, inference.py:131-178).

    Projected BFGS per toy: central-difference gradients (2k points per toy) and a fixed set of trial steps
    (len(line_search) points per toy), each evaluated for all unconverged toys by ONE batch_toys(toy_index=...)
    pass; box bounds as make_objective reports them (rates >= 0, shape parameters inside their anchor range).

    :returns: (dict name -> array [T] of best-fit values, max log likelihood [T], dict(converged=[T], iterations=int,
              evaluations=int))"""
    T = lf.n_toys
    if T == 0:
        raise ValueError("bestfit_toys needs lf.set_toy_data(...) first")
    guess = guess or {}
    names, start, lo, hi = [], [], [], []
    for source_name in lf.rate_parameters.keys():
        key = source_name + _RATE_SUFFIX
        if key not in kwargs:
            names.append(key); start.append(guess.get(key, 1.0)); lo.append(0.0); hi.append(np.inf)
    for setting, (_, _, base_value) in lf.shape_parameters.items():
        if setting in kwargs:
            continue
        s0 = guess.get(setting)
        if s0 is None:
            s0 = lf.pdf_base_config.get(setting)
            if not isinstance(s0, (int, float)):
                s0 = base_value
        b = lf.get_bounds(setting)
        names.append(setting); start.append(s0); lo.append(b[0]); hi.append(b[1])
    if not names:
        raise NoOpimizationNecessary("There are no parameters to fit, no optimization is necessary")
    k = len(names)
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    fixed = list(kwargs.keys())
    columns = names + fixed
    n_eval = [0]

    def minus_ll(points, toy_index):
        cols = np.empty((len(points), k + len(fixed)))
        cols[:, :k] = points
        for j, key in enumerate(fixed):
            cols[:, k + j] = kwargs[key]
        n_eval[0] += len(points)
        return -lf.batch_toys(cols, columns, livetime_days=livetime_days, toy_index=toy_index)

    x0 = np.empty((T, k))
    for j in range(k):
        x0[:, j] = np.broadcast_to(np.asarray(start[j], dtype=np.float64), (T,))
    x, f, converged, it = _lockstep_bfgs(minus_ll, x0, lo, hi, max_iter=max_iter, gtol=gtol, ftol=ftol, fd_step=fd_step,
                                         line_search=line_search)
    result = OrderedDict((n, x[:, j].copy()) for j, n in enumerate(names))
    return result, -f, dict(converged=converged, iterations=it, evaluations=n_eval[0])


def profile_scan(lf, target, values, guess=None, livetime_days=None, max_iter=60, **kwargs):
    """Profile likelihood of ONE dataset over a grid of hypotheses for `target`: the conditional fits of all
    hypotheses (every other free parameter floated) advance in lock step, each iteration being two lf.batch passes
    (the batched form of the loop in one_parameter_interval / plot_likelihood_ratio, inference.py:332-443).

    :returns: (max log likelihood per hypothesis [H], dict name -> conditional best-fit values [H])"""
    values = np.asarray(values, dtype=np.float64).reshape(-1)
    H = len(values)
    guess = guess or {}
    fixed_kwargs = dict(kwargs)
    fixed_kwargs[target] = 0.0                                       # the target is fixed per fit, not floated
    names, start, lo, hi = [], [], [], []
    for source_name in lf.rate_parameters.keys():
        key = source_name + _RATE_SUFFIX
        if key not in fixed_kwargs:
            names.append(key); start.append(guess.get(key, 1.0)); lo.append(0.0); hi.append(np.inf)
    for setting, (_, _, base_value) in lf.shape_parameters.items():
        if setting in fixed_kwargs:
            continue
        s0 = guess.get(setting)
        if s0 is None:
            s0 = lf.pdf_base_config.get(setting)
            if not isinstance(s0, (int, float)):
                s0 = base_value
        b = lf.get_bounds(setting)
        names.append(setting); start.append(s0); lo.append(b[0]); hi.append(b[1])
    fixed = list(kwargs.keys())
    columns = names + [target] + fixed
    k = len(names)

    def minus_ll(points, fit_index):
        cols = np.empty((len(points), k + 1 + len(fixed)))
        cols[:, :k] = points
        cols[:, k] = values[fit_index]
        for j, key in enumerate(fixed):
            cols[:, k + 1 + j] = kwargs[key]
        return -lf.batch(cols, columns, livetime_days=livetime_days)

    if k == 0:
        return -minus_ll(np.zeros((H, 0)), np.arange(H)), OrderedDict()
    x0 = np.tile(np.asarray(start, dtype=np.float64), (H, 1))
    x, f, _, _ = _lockstep_bfgs(minus_ll, x0, np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64),
                                max_iter=max_iter)
    return -f, OrderedDict((n, x[:, j].copy()) for j, n in enumerate(names))


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


def one_parameter_interval_scan(lf, target, bound, confidence_level=0.9, kind='upper', n_grid=17, refinements=3,
                                t_ppf=None, **kwargs):
    """one_parameter_interval with the conditional fits of a whole grid of hypotheses advanced in lock step
    (profile_scan) instead of one bestfit per brentq step: the crossing of the test statistic with its critical value
    is bracketed on n_grid hypotheses, the bracket is re-scanned `refinements` times and the crossing interpolated
    linearly.  Same conventions as one_parameter_interval (inference.py:332-389): Wilks' chi2(1) unless t_ppf is given;
    kind 'upper' / 'lower' / 'central'; bound = search limit (2-tuple for 'central')."""
    if target is None:
        target = lf.source_list[-1] + _RATE_SUFFIX
    # global best fit: the lock-step fitter on a single 'hypothesis-free' fit (target floats)
    names_free = [n for n in lf.parameter_names() if n not in kwargs]
    x_best, f_best = _global_fit(lf, names_free, kwargs)
    global_best, max_ll = x_best[target], f_best

    def crossing(lo, hi, quantile, rising):
        """hypothesis in [lo, hi] where 2 (max_ll - ll_cond(h)) - critical(h) changes sign."""
        for _ in range(refinements + 1):
            grid = np.linspace(lo, hi, n_grid)
            prof, _ = profile_scan(lf, target, grid, **kwargs)
            crit = np.array([stats.norm.ppf(quantile) ** 2 if t_ppf is None else t_ppf(h, quantile) for h in grid])
            t = 2 * (max_ll - prof) - crit
            sign = t > 0
            flips = np.flatnonzero(sign[1:] != sign[:-1])
            if not len(flips):
                raise OptimizationFailed("the test statistic does not cross its critical value inside the bound")
            i = flips[0] if rising else flips[-1]
            lo, hi = grid[i], grid[i + 1]
            t_lo, t_hi = t[i], t[i + 1]
        return lo + (hi - lo) * (0.0 - t_lo) / (t_hi - t_lo)

    if kind == 'upper':
        return crossing(global_best, bound, confidence_level, True)
    if kind == 'lower':
        return crossing(bound, global_best, 1 - confidence_level, False)
    if kind == 'central':
        return (crossing(bound[0], global_best, (1 - confidence_level) / 2, False),
                crossing(global_best, bound[1], 1 - (1 - confidence_level) / 2, True))
    raise ValueError("kind must be 'upper', 'lower' or 'central'")


def _global_fit(lf, names_free, fixed):
    """Best fit of one dataset over names_free with the lock-step BFGS core (one fit): ({name: value}, max logL)."""
    lo, hi, start = [], [], []
    for n in names_free:
        if n.endswith(_RATE_SUFFIX):
            lo.append(0.0); hi.append(np.inf); start.append(1.0)
        else:
            b = lf.get_bounds(n)
            s0 = lf.pdf_base_config.get(n)
            if not isinstance(s0, (int, float)):
                s0 = lf.shape_parameters[n][2]
            lo.append(b[0]); hi.append(b[1]); start.append(s0)
    fixed_names = list(fixed.keys())
    columns = names_free + fixed_names

    def minus_ll(points, fit_index):
        cols = np.empty((len(points), len(columns)))
        cols[:, :len(names_free)] = points
        for j, key in enumerate(fixed_names):
            cols[:, len(names_free) + j] = fixed[key]
        return -lf.batch(cols, columns)

    x, f, _, _ = _lockstep_bfgs(minus_ll, np.asarray(start, dtype=np.float64)[None, :],
                                np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64))
    return dict(zip(names_free, x[0])), -f[0]


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
