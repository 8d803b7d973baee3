"""Oracle: whole-path CPU pipelines at array level (test infrastructure, see oracle/__init__.py).

These classes do, with the same third-party calls and in the same order, what the reference
does between `set_data` and `lf(**params)` -- so they double as the timed CPU baseline
(`bench.py`'s cpu_baseline / `--impl reference`, kind "port"):

  set_data  : per anchor model, per source, one pdf lookup of all events
              (likelihood.py:557-560 -> model.py:97-99 -> source.py:219-246), stored in a dense
              float64 tensor [n1..nD, S, N] (pdf_morphers.py:59-65), then ONE
              RegularGridInterpolator over it (pdf_morphers.py:67)
  __call__  : bounds test, two RGI calls (mus, ps), rate scaling, unphysical check,
              extended_loglikelihood (likelihood.py:339-427, :678-690)
"""
import numpy as np
from scipy import stats
from scipy.interpolate import RegularGridInterpolator

from . import binned as _binned
from . import hist as _hist
from . import unbinned as _unbinned


class UnbinnedOracle(object):
    def __init__(self, axes, mus_anchor, outlier_likelihood=1e-12, allow_negative=None):
        self.axes = [np.asarray(a, dtype=float) for a in axes]
        self.mus_anchor = np.asarray(mus_anchor, dtype=float)
        self.outlier_likelihood = outlier_likelihood
        self.allow_negative = allow_negative
        self.ps_anchor = None
        if len(self.axes):
            self._mus_itp = RegularGridInterpolator(self.axes, self.mus_anchor)

    # -- set_data ------------------------------------------------------------------------------
    def set_ps(self, ps_anchor):
        self.ps_anchor = np.asarray(ps_anchor, dtype=float)
        if len(self.axes):
            self._ps_itp = RegularGridInterpolator(self.axes, self.ps_anchor)
        return self

    def set_data_from_templates(self, templates, edges_list, coords, method='linear'):
        """templates [n1..nD, S, *bins] of pdf densities; coords = list of D_space arrays of N."""
        templates = np.asarray(templates, dtype=float)
        d = len(self.axes)
        grid_shape = templates.shape[:d]
        n_sources = templates.shape[d]
        n = len(coords[0])
        ps = np.zeros(list(grid_shape) + [n_sources, n])
        look = _hist.lookup_linear if method == 'linear' else _hist.lookup_piecewise
        for g in np.ndindex(*grid_shape):
            ps[g] = np.vstack([look(templates[g + (s,)], edges_list, coords) for s in range(n_sources)])
        return self.set_ps(ps)

    # -- evaluation ----------------------------------------------------------------------------
    def __call__(self, zs, rate_multipliers, livetime_scale=None, full_output=False):
        zs = np.asarray(zs, dtype=float)
        for a, z in zip(self.axes, zs):
            if not a[0] <= z <= a[-1]:
                return -float('inf')
        if len(self.axes):
            mus = self._mus_itp(zs)[0]
            ps = self._ps_itp(zs)[0]
        else:
            mus = self.mus_anchor.copy()
            ps = self.ps_anchor
        mus = _unbinned.scale_mus(mus, rate_multipliers, livetime_scale)
        if _unbinned.rates_unphysical(mus, self.allow_negative):
            return -float('inf')
        ll = _unbinned.extended_loglikelihood(mus, ps, self.outlier_likelihood)
        if full_output:
            return ll, mus, ps
        return ll

    def batch(self, zs_array, rate_multiplier_array):
        return np.array([self(z, m) for z, m in zip(zs_array, rate_multiplier_array)])


class SourcewiseUnbinnedOracle(object):
    """Source-wise interpolation (likelihood.py:113-145,152-169,210-240,534-555): one
    RegularGridInterpolator per source over the shape parameters that source depends on
    (pdf_morphers.py:45-70 with the source's own shape_parameters); sources without shape
    parameters keep their base values.

    axes: anchor axes of ALL shape parameters (bounds test, likelihood.py:345-347);
    source_dims[s]: indices of the parameters source s depends on (ascending);
    mus_sub[s] / ps_sub[s]: arrays shaped [sub-grid...] / [sub-grid..., N] (scalars / [N] for no parameters)."""

    def __init__(self, axes, source_dims, mus_sub, outlier_likelihood=1e-12, allow_negative=None):
        self.axes = [np.asarray(a, dtype=float) for a in axes]
        self.source_dims = [list(d) for d in source_dims]
        self.outlier_likelihood = outlier_likelihood
        self.allow_negative = allow_negative
        self.mus_sub = [np.asarray(m, dtype=float) for m in mus_sub]
        self._mus_itp = [RegularGridInterpolator([self.axes[d] for d in dims], m.reshape(m.shape + (1,)))
                         if len(dims) else None for dims, m in zip(self.source_dims, self.mus_sub)]
        self._ps_itp = None

    def set_ps(self, ps_sub):
        self.ps_sub = [np.asarray(p, dtype=float) for p in ps_sub]
        self._ps_itp = [RegularGridInterpolator([self.axes[d] for d in dims], p) if len(dims) else None
                        for dims, p in zip(self.source_dims, self.ps_sub)]
        return self

    def __call__(self, zs, rate_multipliers, livetime_scale=None, full_output=False):
        zs = np.asarray(zs, dtype=float)
        for a, z in zip(self.axes, zs):
            if not a[0] <= z <= a[-1]:
                return -float('inf')
        mus, ps = [], []
        for s, dims in enumerate(self.source_dims):
            if len(dims):
                these = np.asarray([zs[d] for d in dims])
                mus.append(self._mus_itp[s](these)[0][0])          # likelihood.py:230 (extra_dims=[1], then [0])
                ps.append(self._ps_itp[s](these)[0])               # likelihood.py:550
            else:
                mus.append(float(self.mus_sub[s]))
                ps.append(self.ps_sub[s])
        mus = _unbinned.scale_mus(np.array(mus), rate_multipliers, livetime_scale)
        ps = np.array(ps)
        if _unbinned.rates_unphysical(mus, self.allow_negative):
            return -float('inf')
        ll = _unbinned.extended_loglikelihood(mus, ps, self.outlier_likelihood)
        if full_output:
            return ll, mus, ps
        return ll

    def batch(self, zs_array, rate_multiplier_array):
        return np.array([self(z, m) for z, m in zip(zs_array, rate_multiplier_array)])


class BinnedOracle(object):
    """likelihood.py:576-675.  pmf_anchor / n_model_anchor: [n1..nD, S, *bins]."""

    def __init__(self, axes, mus_anchor, pmf_anchor, n_model_anchor=None, bb_source=None):
        self.axes = [np.asarray(a, dtype=float) for a in axes]
        self.mus_anchor = np.asarray(mus_anchor, dtype=float)
        self.pmf_anchor = np.asarray(pmf_anchor, dtype=float)
        self.n_model_anchor = None if n_model_anchor is None else np.asarray(n_model_anchor, dtype=float)
        self.bb_source = bb_source
        if len(self.axes):
            self._mus_itp = RegularGridInterpolator(self.axes, self.mus_anchor)
            self._pmf_itp = RegularGridInterpolator(self.axes, self.pmf_anchor)
            if self.n_model_anchor is not None:
                self._nm_itp = RegularGridInterpolator(self.axes, self.n_model_anchor)
        self.observed = None

    def set_observed(self, observed):
        self.observed = np.asarray(observed, dtype=float)
        return self

    def set_data(self, edges_list, coords):
        return self.set_observed(_hist.histogramdd(edges_list, coords))

    def __call__(self, zs, rate_multipliers, livetime_scale=None, full_output=False):
        zs = np.asarray(zs, dtype=float)
        for a, z in zip(self.axes, zs):
            if not a[0] <= z <= a[-1]:
                return -float('inf')
        if len(self.axes):
            mus = self._mus_itp(zs)[0]
            pmfs = self._pmf_itp(zs)[0]
            nm = self._nm_itp(zs)[0] if self.n_model_anchor is not None else None
        else:
            mus = self.mus_anchor.copy()
            pmfs = self.pmf_anchor
            nm = self.n_model_anchor
        mus = _unbinned.scale_mus(mus, rate_multipliers, livetime_scale)
        if _unbinned.rates_unphysical(mus, None):
            return -float('inf')
        if self.bb_source is not None:
            mus, pmfs = _binned.adjust_expectations_bb(mus, pmfs, nm, self.observed, self.bb_source)
        else:
            mus, pmfs = mus.copy(), pmfs.copy()
        ll = _binned.binned_loglikelihood(mus, pmfs, self.observed)
        if full_output:
            return ll, mus, pmfs
        return ll

    def batch(self, zs_array, rate_multiplier_array):
        return np.array([self(z, m) for z, m in zip(zs_array, rate_multiplier_array)])


def toy_loglikelihoods(axes, mus_anchor, templates, edges_list, coords, offsets, zs_array, rate_multiplier_array,
                       method='linear', outlier_likelihood=1e-12):
    """Many datasets, one parameter point each: toy t = events offsets[t]:offsets[t+1] of `coords`, evaluated at
    (zs_array[t], rate_multiplier_array[t]).  The reference has no such call; this is literally its loop
    `lf.set_data(toy_t); lf(**theta_t)` (likelihood.py:531-562, :318-427) over the toys."""
    out = np.empty(len(offsets) - 1)
    for t in range(len(offsets) - 1):
        sl = slice(int(offsets[t]), int(offsets[t + 1]))
        orc = UnbinnedOracle(axes, mus_anchor, outlier_likelihood=outlier_likelihood)
        orc.set_data_from_templates(templates, edges_list, [np.asarray(c)[sl] for c in coords], method)
        out[t] = orc(zs_array[t], rate_multiplier_array[t])
    return out
