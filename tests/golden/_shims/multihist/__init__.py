"""TEST-ONLY stand-in for the `multihist` package (not installed here, no network).

Only used by tests/golden/make_golden.py to import the *unmodified* reference from
/root/reference inside the build container.  It is never imported by blueice_b200.
Semantics are restated from memory of multihist 0.6.x for exactly the call sites the
reference uses: Histdd(bins=, axis_names=), add, histogram, n,
bin_centers, lookup, similar_blank_hist, __mul__, get_random.
`lookup` and `get_random` are "parity unpinned" (no reference test reaches them).
"""
import numpy as np


class Histdd(object):
    def __init__(self, *data, bins=10, axis_names=None, weights=None):
        self.axis_names = axis_names
        self.bin_edges = [np.asarray(b, dtype=float) for b in bins]
        self.dimensions = len(self.bin_edges)
        self.histogram = np.zeros([len(b) - 1 for b in self.bin_edges], dtype=float)
        if len(data):
            self.add(*data, weights=weights)

    def add(self, *data, weights=None):
        sample = np.array(data).T
        if sample.size == 0:
            return
        h, _ = np.histogramdd(sample, bins=self.bin_edges, weights=weights)
        self.histogram = self.histogram + h

    @property
    def n(self):
        return self.histogram.sum()

    def bin_centers(self, axis=None):
        if axis is None:
            return [self.bin_centers(i) for i in range(self.dimensions)]
        e = self.bin_edges[axis]
        return 0.5 * (e[1:] + e[:-1])

    def similar_blank_hist(self):
        h = Histdd(bins=self.bin_edges, axis_names=self.axis_names)
        return h

    def lookup(self, *coordinate_arrays):
        idx = []
        for e, x in zip(self.bin_edges, coordinate_arrays):
            i = np.searchsorted(e, x) - 1
            idx.append(np.clip(i, 0, len(e) - 2))
        return self.histogram[tuple(idx)]

    def __mul__(self, other):
        h = self.similar_blank_hist()
        h.histogram = self.histogram * other
        return h

    def get_random(self, size=10):
        size = int(size)
        flat = self.histogram.ravel()
        cdf = np.cumsum(flat)
        cdf = cdf / cdf[-1]
        which = np.searchsorted(cdf, np.random.rand(size))
        which = np.clip(which, 0, len(flat) - 1)
        multi = np.unravel_index(which, self.histogram.shape)
        out = np.zeros((size, self.dimensions))
        for d, e in enumerate(self.bin_edges):
            lo = e[multi[d]]
            hi = e[multi[d] + 1]
            out[:, d] = lo + np.random.rand(size) * (hi - lo)
        return out
