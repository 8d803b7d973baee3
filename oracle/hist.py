"""Oracle: histogram semantics (test infrastructure, see oracle/__init__.py).

 * histogramdd binning  = multihist.Histdd.add -> np.histogramdd (likelihood.py:604-609)
 * piecewise pdf lookup = multihist.Histdd.lookup (source.py:242-243)   [PARITY UNPINNED]
 * linear pdf lookup    = RegularGridInterpolator over bin centres with clipping (source.py:225-240)
"""
import numpy as np
from scipy.interpolate import RegularGridInterpolator

from . import morph


def bin_centers(edges):
    edges = np.asarray(edges, dtype=float)
    return 0.5 * (edges[1:] + edges[:-1])


def histogramdd_indices(edges_list, coords):
    """Per-event flat bin index (C order) or -1 when the event is dropped.

    np.histogramdd rule (pinned in tests/test_oracle_pins.py): bin = searchsorted(edges, x,
    side='right') - 1; x == last edge -> last bin; x < first edge, x > last edge or NaN -> dropped.
    """
    coords = [np.asarray(c, dtype=float) for c in coords]
    n = len(coords[0]) if len(coords) else 0
    flat = np.zeros(n, dtype=np.int64)
    ok = np.ones(n, dtype=bool)
    for e, x in zip(edges_list, coords):
        e = np.asarray(e, dtype=float)
        nb = len(e) - 1
        i = np.searchsorted(e, x, side='right') - 1
        i = np.where(x == e[-1], nb - 1, i)
        good = (i >= 0) & (i < nb) & ~np.isnan(x)
        ok &= good
        flat = flat * nb + np.where(good, i, 0)
    return np.where(ok, flat, -1)


def histogramdd(edges_list, coords):
    """float64 counts, the reference's data_events_per_bin.histogram (likelihood.py:608-609)."""
    shape = [len(e) - 1 for e in edges_list]
    sample = np.array([np.asarray(c, dtype=float) for c in coords]).T
    if sample.size == 0:
        return np.zeros(shape)
    h, _ = np.histogramdd(sample, bins=[np.asarray(e, dtype=float) for e in edges_list])
    return h


def lookup_piecewise_indices(edges_list, coords):
    """multihist Histdd.lookup index rule (from memory; unpinned):
    idx_d = clip(searchsorted(edges_d, x_d, side='left') - 1, 0, nbins_d - 1)."""
    idx = []
    for e, x in zip(edges_list, coords):
        e = np.asarray(e, dtype=float)
        i = np.searchsorted(e, np.asarray(x, dtype=float)) - 1
        idx.append(np.clip(i, 0, len(e) - 2))
    return idx


def lookup_piecewise(hist, edges_list, coords):
    return np.asarray(hist)[tuple(lookup_piecewise_indices(edges_list, coords))]


def lookup_linear(hist, edges_list, coords):
    """source.py:225-240: clip to the bin-centre range, then RGI over bin centres."""
    hist = np.array(hist, dtype=float)
    centers = [bin_centers(e) for e in edges_list]
    itp = RegularGridInterpolator(centers, hist)
    clipped = [np.clip(np.asarray(x, dtype=float), c.min(), c.max()) for x, c in zip(coords, centers)]
    return itp(np.transpose(clipped))


def lookup_linear_explicit(hist, edges_list, coords):
    """Explicit restatement of lookup_linear with SciPy's operation orders:
    2-D histograms take the Cython fast path `evaluate_linear_2d` (_rgi.py:448-462):
        r = 0; r += V00*(1-y0)*(1-y1); r += V01*(1-y0)*y1; r += V10*y0*(1-y1); r += V11*y0*y1
    every other dimensionality takes the generic corner loop (oracle/morph.py)."""
    hist = np.asarray(hist, dtype=float)
    centers = [bin_centers(e) for e in edges_list]
    xs = [np.clip(np.asarray(x, dtype=float), c.min(), c.max()) for x, c in zip(coords, centers)]
    cells = [morph.find_cells(c, x) for c, x in zip(centers, xs)]
    if hist.ndim == 2:
        (i0, y0), (i1, y1) = cells
        r = np.zeros(len(xs[0]))
        r = r + hist[i0, i1] * (1 - y0) * (1 - y1)
        r = r + hist[i0, i1 + 1] * (1 - y0) * y1
        r = r + hist[i0 + 1, i1] * y0 * (1 - y1)
        r = r + hist[i0 + 1, i1 + 1] * y0 * y1
        return r
    import itertools
    acc = np.zeros(len(xs[0]))
    for bits in itertools.product((0, 1), repeat=hist.ndim):
        w = np.ones(len(xs[0]))
        idx = []
        for (i, y), b in zip(cells, bits):
            w = w * (y if b else 1 - y)
            idx.append(i + b)
        acc = acc + hist[tuple(idx)] * w
    return acc
