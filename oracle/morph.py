"""Oracle: anchor-grid multilinear morphing (test infrastructure, see oracle/__init__.py).

Follows blueice/pdf_morphers.py:45-80 (GridInterpolator) and the SciPy code it calls:
scipy/interpolate/_rgi.py:375-480 (__call__), :520-549 (_evaluate_linear) and the Cython
`find_indices`.  The explicit rules below are verified bit-for-bit against SciPy in
tests/test_oracle_pins.py.
"""
import itertools

import numpy as np
from scipy.interpolate import RegularGridInterpolator


def anchor_axes(anchor_z_lists):
    """Per shape parameter: sorted array of anchor z values (pdf_morphers.py:48-49)."""
    return [np.array(sorted(zs), dtype=float) for zs in anchor_z_lists]


def anchor_points(axes):
    """All anchor points in C order, first axis slowest (pdf_morphers.py:53-54,72-80)."""
    return list(itertools.product(*[list(a) for a in axes]))


def find_cell(axis, z):
    """Cell index and normalised distance along one axis.

    Rule pinned against scipy's find_indices:
        i = clip(searchsorted(axis, z, side='right') - 1, 0, n - 2);  y = (z - a[i]) / (a[i+1] - a[i])
    so z exactly on an interior anchor selects the upper cell with y == 0, and z on the last
    anchor selects cell n-2 with y == 1.  A one-point axis gives i = -1, y = 0
    (both corners alias value 0).
    """
    axis = np.asarray(axis, dtype=float)
    n = len(axis)
    if n == 1:
        return -1, 0.0
    i = int(np.clip(np.searchsorted(axis, z, side='right') - 1, 0, n - 2))
    y = (z - axis[i]) / (axis[i + 1] - axis[i])
    return i, float(y)


def find_cells(axis, z):
    """Vectorised find_cell for an array of z."""
    axis = np.asarray(axis, dtype=float)
    z = np.asarray(z, dtype=float)
    n = len(axis)
    if n == 1:
        return np.full(z.shape, -1, dtype=np.int64), np.zeros(z.shape)
    i = np.clip(np.searchsorted(axis, z, side='right') - 1, 0, n - 2).astype(np.int64)
    y = (z - axis[i]) / (axis[i + 1] - axis[i])
    return i, y


def corner_table(axes, zs):
    """Hypercube corners (first dim slowest) with weights in SciPy's operation order.

    Returns (corner_indices [C, D], weights [C]) following _rgi.py:538-547:
        weight = 1.; for each dim: weight = weight * (1 - y_d  or  y_d)
    """
    cells = [find_cell(a, z) for a, z in zip(axes, zs)]
    corners, weights = [], []
    for bits in itertools.product((0, 1), repeat=len(axes)):
        w = 1.0
        idx = []
        for (i, y), b in zip(cells, bits):
            w = w * (y if b else 1 - y)
            idx.append(i + b)
        corners.append(idx)
        weights.append(w)
    return np.array(corners, dtype=np.int64).reshape(len(corners), len(axes)), np.array(weights)


def morph_explicit(axes, values, zs):
    """Explicit restatement of itp(zs)[0]: value = value + values[corner] * weight, corner order
    = itertools.product over dims, first dim slowest (_rgi.py:544-549)."""
    values = np.asarray(values, dtype=float)
    d = len(axes)
    if d == 0:
        return values.copy()
    corners, weights = corner_table(axes, zs)
    acc = np.zeros(values.shape[d:])
    for idx, w in zip(corners, weights):
        acc = acc + values[tuple(idx)] * w
    return acc


def morph_rgi(axes, values):
    """The reference's own construction: RegularGridInterpolator closure (pdf_morphers.py:67-70)."""
    itp = RegularGridInterpolator(axes, values)
    return lambda zs: itp(np.asarray(zs, dtype=float))[0]
