"""Minimal N-dimensional histogram used by the sources (host side).

The reference depends on `multihist.Histdd` (not vendored, not installable here).  This class offers
exactly the surface the hot path's callers use with the same semantics:
  add            = np.histogramdd accumulate            (likelihood.py:609, source.py:298)
  histogram / n / bin_edges / bin_centers / similar_blank_hist / __mul__
  lookup         = hist[clip(searchsorted(edges, x, 'left') - 1, 0, nbins - 1)]   (source.py:243)
  get_random     = bin chosen in proportion to its content, uniform inside the bin (source.py:254)
The device versions of `add` and `lookup` are bi_histogramdd / bi_hist_lookup.
"""
import numpy as np

__all__ = ['Histdd']


class Histdd(object):
    def __init__(self, *data, bins=10, axis_names=None, weights=None):
        self.bin_edges = [np.asarray(b, dtype=float) for b in bins]
        self.axis_names = axis_names
        self.dimensions = len(self.bin_edges)
        self.histogram = np.zeros([len(e) - 1 for e in self.bin_edges], dtype=float)
        if data:
            self.add(*data, weights=weights)

    @classmethod
    def from_histogram(cls, histogram, bin_edges, axis_names=None):
        self = cls(bins=bin_edges, axis_names=axis_names)
        self.histogram = np.asarray(histogram, dtype=float)
        assert self.histogram.shape == tuple(len(e) - 1 for e in self.bin_edges)
        return self

    def add(self, *coordinate_arrays, weights=None):
        sample = np.array(coordinate_arrays).T
        if sample.size:
            counts, _ = np.histogramdd(sample, bins=self.bin_edges, weights=weights)
            self.histogram = self.histogram + counts

    @property
    def n(self):
        return self.histogram.sum()

    def bin_centers(self, axis=None):
        if axis is None:
            return [self.bin_centers(i) for i in range(self.dimensions)]
        e = self.bin_edges[axis]
        return 0.5 * (e[1:] + e[:-1])

    def bin_volumes(self):
        vol = np.ones(1)
        for e in self.bin_edges:
            vol = np.multiply.outer(vol, np.diff(e))
        return vol.reshape(self.histogram.shape)

    def similar_blank_hist(self):
        return Histdd(bins=self.bin_edges, axis_names=self.axis_names)

    def lookup(self, *coordinate_arrays):
        index = []
        for e, x in zip(self.bin_edges, coordinate_arrays):
            index.append(np.clip(np.searchsorted(e, x) - 1, 0, len(e) - 2))
        return self.histogram[tuple(index)]

    def __mul__(self, other):
        out = self.similar_blank_hist()
        out.histogram = self.histogram * other
        return out

    __rmul__ = __mul__

    def get_random(self, size=10):
        size = int(size)
        flat = self.histogram.ravel()
        cdf = np.cumsum(flat)
        picks = np.searchsorted(cdf / cdf[-1], np.random.rand(size))
        picks = np.minimum(picks, len(flat) - 1)
        multi = np.unravel_index(picks, self.histogram.shape)
        out = np.empty((size, self.dimensions))
        for d, e in enumerate(self.bin_edges):
            lo, hi = e[multi[d]], e[multi[d] + 1]
            out[:, d] = lo + np.random.rand(size) * (hi - lo)
        return out
