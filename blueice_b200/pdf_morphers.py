"""Morphers: interpolate functions of anchor models over the shape-parameter space.

Mirror of the reference seam blueice/pdf_morphers.py:15-80 (`Morpher`, `GridInterpolator`, the
`MORPHERS` registry selected by likelihood_config['morpher']).  The multilinear interpolation that
the reference delegates to scipy's RegularGridInterpolator (pdf_morphers.py:67-70) runs on the GPU:
cells, fractions and corner weights by bi_point_setup (K1), the weighted corner sum in SciPy's own
operation order by bi_unbinned_ps -- bit-identical results.

RadialInterpolator (pdf_morphers.py:83-193, "highly experimental") is out of scope (SURVEY.md
section 2, row 9) and is not registered.
"""
import itertools

import numpy as np

from .exceptions import NoShapeParameters
from .utils import arrays_to_grid

__all__ = ['Morpher', 'GridInterpolator', 'MORPHERS']


class Morpher(object):
    """Interface: which anchor models are needed, and how to interpolate between them."""

    def __init__(self, config, shape_parameters):
        self.config = config
        self.shape_parameters = shape_parameters
        if not len(self.shape_parameters):
            raise NoShapeParameters("Attempt to initialize a morpher without shape parameters")

    def get_anchor_points(self, bounds, n_models=None):
        """List of anchor z-tuples at which models have to be computed."""
        raise NotImplementedError

    def make_interpolator(self, f, extra_dims, anchor_models):
        """callable(zs) -> array of shape extra_dims interpolating f(model) between the anchors."""
        raise NotImplementedError


class DeviceGridFunction(object):
    """values[n1..nD, *extra_dims] resident in HBM, evaluated at one point per call."""

    def __init__(self, axes, values, extra_dims):
        from .engine import MorphGrid, UnbinnedEngine
        self.extra_dims = list(extra_dims)
        self.grid = MorphGrid(axes)
        flat = np.asarray(values, dtype=np.float64).reshape(self.grid.n_anchors, 1, -1)
        self.engine = UnbinnedEngine(self.grid, np.zeros((self.grid.n_anchors, 1)))
        self.engine.set_ps_anchor(flat)

    def __call__(self, zs):
        zs = np.asarray(zs, dtype=np.float64).reshape(-1)
        if len(zs) != self.grid.n_dims:
            raise ValueError("The requested sample points xi have dimension %d but this "
                             "interpolator has dimension %d" % (len(zs), self.grid.n_dims))
        ok = True
        for d, (a, z) in enumerate(zip(self.grid.axes, zs)):
            if not (a[0] <= z <= a[-1]):          # scipy bounds_error=True (also rejects NaN)
                raise ValueError("One of the requested xi is out of bounds in dimension %d" % d)
        _, ps = self.engine.ps(zs, [1.0])
        return ps[0].reshape(self.extra_dims) if self.extra_dims else ps[0].reshape(())


class GridInterpolator(Morpher):
    """Multilinear interpolation on the regular grid spanned by the sorted anchor values."""

    def __init__(self, config, shape_parameters):
        super().__init__(config, shape_parameters)
        # one sorted axis per shape parameter, in insertion order (pdf_morphers.py:48-49)
        self.anchor_z_arrays = [np.array(sorted(anchors.keys()))
                                for _, (anchors, _, _) in shape_parameters.items()]
        self.anchor_z_grid = arrays_to_grid(self.anchor_z_arrays)

    def _anchor_grid_iterator(self):
        """Yield (grid index list, z tuple) in C order, first parameter slowest (pdf_morphers.py:72-80)."""
        for index in itertools.product(*[range(len(a)) for a in self.anchor_z_arrays]):
            yield list(index), tuple(a[i] for a, i in zip(self.anchor_z_arrays, index))

    def get_anchor_points(self, bounds, n_models=None):
        return [zs for _, zs in self._anchor_grid_iterator()]

    def anchor_tensor(self, f, extra_dims, anchor_models):
        """Dense host tensor [n1..nD, *extra_dims] of f(model) at every anchor (pdf_morphers.py:59-65)."""
        scores = np.zeros([len(a) for a in self.anchor_z_arrays] + list(extra_dims))
        for index, zs in self._anchor_grid_iterator():
            scores[tuple(index)] = f(anchor_models[tuple(zs)])
        return scores

    def make_interpolator(self, f, extra_dims, anchor_models):
        return DeviceGridFunction(self.anchor_z_arrays, self.anchor_tensor(f, extra_dims, anchor_models),
                                  extra_dims)


MORPHERS = {'GridInterpolator': GridInterpolator}
