"""Oracle: binned Poisson likelihood with Beeston-Barlow (test infrastructure).

Follows blueice/likelihood.py:618-675 (adjust_expectations, _compute_likelihood) and
:693-712 (beeston_barlow_root1/2).  poisson_logpmf restates scipy.stats.poisson(lam).logpmf(k)
(scipy/stats/_discrete_distns.py:999-1001 + the argument checks of
_distn_infrastructure.py:3544-3580) and is pinned against SciPy in tests/test_oracle_pins.py.
"""
import numpy as np
from scipy import special, stats


def beeston_barlow_root1(a, p, U, d):
    """likelihood.py:693-700 (the root the reference asserts to be <= 0)."""
    with np.errstate(all='ignore'):
        return ((-U*p - U + a*p + d*p -
                 np.sqrt(U**2*p**2 + 2*U**2*p + U**2 + 2*U*a*p**2 + 2*U*a*p -
                         2*U*d*p**2 - 2*U*d*p + a**2*p**2 + 2*a*d*p**2 + d**2*p**2))/(2*p*(p + 1)))


def beeston_barlow_root2(a, p, U, d):
    """likelihood.py:703-708."""
    with np.errstate(all='ignore'):
        return ((-U*p - U + a*p + d*p +
                 np.sqrt(U**2*p**2 + 2*U**2*p + U**2 + 2*U*a*p**2 + 2*U*a*p -
                         2*U*d*p**2 - 2*U*d*p + a**2*p**2 + 2*a*d*p**2 + d**2*p**2))/(2*p*(p + 1)))


def adjust_expectations_bb(mus, pmfs, n_model_events, observed, source_i):
    """likelihood.py:620-660 with model_statistical_uncertainty_handling == 'bb_single'."""
    mus = np.array(mus, dtype=float)
    pmfs = np.array(pmfs, dtype=float)
    assert pmfs.shape == n_model_events.shape
    counts = pmfs.copy()
    for i, mu in enumerate(mus):
        counts[i] *= (mu if i != source_i else 0.)
    u_bins = np.sum(counts, axis=0)
    a_bins = n_model_events[source_i]
    with np.errstate(all='ignore'):
        p_cal = mus[source_i] / n_model_events[source_i].sum()
        w_cal = pmfs[source_i] / a_bins * n_model_events[source_i].sum()
        A1 = beeston_barlow_root1(a_bins, w_cal * p_cal, u_bins, observed)
        A2 = beeston_barlow_root2(a_bins, w_cal * p_cal, u_bins, observed)
        assert np.all(A1 <= 0)
        A_special = (observed + a_bins) / (1. + p_cal)
        A = np.choose(u_bins == 0, [A2, A_special])
        assert np.all(0 <= A)
        pmfs[source_i] = A * w_cal
        pmfs[source_i] /= pmfs[source_i].sum()
        mus[source_i] = (A * w_cal).sum() * p_cal
    return mus, pmfs


def poisson_logpmf(k, lam):
    """Explicit restatement of scipy.stats.poisson(lam).logpmf(k) edge semantics (pinned against
    SciPy in tests/test_oracle_pins.py):
        lam < 0 or NaN lam -> nan;  NaN k -> nan;  k < 0 or non-integer k -> -inf;
        otherwise xlogy(k, lam) - gammaln(k + 1) - lam
    (0 at lam = k = 0, -inf at lam = 0 < k, -inf at lam = inf, k = 0, nan at lam = inf, k > 0)."""
    k = np.asarray(k, dtype=float)
    lam = np.asarray(lam, dtype=float)
    k, lam = np.broadcast_arrays(k, lam)
    with np.errstate(all='ignore'):
        core = special.xlogy(k, lam) - special.gammaln(k + 1) - lam
        out = np.where((k < 0) | (k != np.floor(k)), -np.inf, core)
        out = np.where(np.isnan(k) | ~(lam >= 0), np.nan, out)
    return out


def binned_loglikelihood(mus, pmfs, observed):
    """likelihood.py:662-675 using SciPy exactly like the reference."""
    expected = np.array(pmfs, dtype=float)
    for mu, row in zip(mus, expected):
        row *= mu
    total = np.sum(expected, axis=0)
    with np.errstate(all='ignore'):
        return np.sum(stats.poisson(total).logpmf(observed))
