"""Oracle: unbinned extended log-likelihood (test infrastructure, see oracle/__init__.py).

Follows blueice/likelihood.py:318-427 (LogLikelihoodBase.__call__ numeric tail) and
:678-690 (extended_loglikelihood).
"""
import numpy as np

from . import morph


def extended_loglikelihood(mu, ps, outlier_likelihood=0.0):
    """likelihood.py:678-690, statement for statement."""
    mu = np.asarray(mu, dtype=float)
    with np.errstate(all='ignore'):
        p_events = np.nansum(mu[:, np.newaxis] * ps, axis=0)
        if outlier_likelihood != 0:
            p_events[True ^ (p_events > 0)] = outlier_likelihood
        return -mu.sum() + np.sum(np.log(p_events))


def scale_mus(mus, rate_multipliers, livetime_scale=None, eff_mask=None, effs=None):
    """likelihood.py:366-393: mus[s] *= mult_s; mus *= livetime/base; mus[eff_mask] *= effs."""
    mus = np.array(mus, dtype=float)
    for s, m in enumerate(rate_multipliers):
        mus[s] *= m
    if livetime_scale is not None:
        mus *= livetime_scale
    if eff_mask is not None and np.any(eff_mask):
        mus[np.asarray(eff_mask, dtype=bool)] *= np.asarray(effs, dtype=float)
    return mus


def rates_unphysical(mus, allow_negative=None):
    """likelihood.py:397-415.  True -> the reference returns -inf (or raises in 'error' mode)."""
    mus = np.asarray(mus, dtype=float)
    if allow_negative is None or not any(allow_negative):
        return not np.all((mus >= 0) & (mus < float('inf')))
    if (not any(mus < float('inf'))) or (np.sum(mus) < 0):
        return True
    for mu, ok in zip(mus, allow_negative):
        if not (0 <= mu) and (not ok):
            return True
    return False


def unbinned_ll_point(axes, mus_anchor, ps_anchor, zs, rate_multipliers, outlier_likelihood=1e-12,
                      livetime_scale=None, eff_mask=None, effs=None, allow_negative=None,
                      use_rgi=True, full_output=False):
    """One likelihood evaluation without priors (likelihood.py:339-427).

    axes        list of D sorted anchor axes
    mus_anchor  [n1..nD, S]   expected events at each anchor (likelihood.py:248-251)
    ps_anchor   [n1..nD, S, N] per-event pdf values at each anchor (likelihood.py:557-560)
    """
    zs = np.asarray(zs, dtype=float)
    for a, z in zip(axes, zs):
        if not a[0] <= z <= a[-1]:          # likelihood.py:345-347 (NaN fails -> -inf)
            return (-float('inf'), None, None) if full_output else -float('inf')
    if len(axes):
        if use_rgi:
            mus = morph.morph_rgi(axes, np.asarray(mus_anchor, dtype=float))(zs)
            ps = morph.morph_rgi(axes, np.asarray(ps_anchor, dtype=float))(zs)
        else:
            mus = morph.morph_explicit(axes, mus_anchor, zs)
            ps = morph.morph_explicit(axes, ps_anchor, zs)
    else:
        mus = np.array(mus_anchor, dtype=float)
        ps = np.asarray(ps_anchor, dtype=float)
    mus = scale_mus(mus, rate_multipliers, livetime_scale, eff_mask, effs)
    if rates_unphysical(mus, allow_negative):
        return (-float('inf'), mus, ps) if full_output else -float('inf')
    ll = extended_loglikelihood(mus, ps, outlier_likelihood)
    return (ll, mus, ps) if full_output else ll
