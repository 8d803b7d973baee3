"""Small host utilities shared by the model / source / likelihood mirrors.

Behaviour follows blueice/utils.py (combine_dicts :27-40, pickle helpers :65-77, hashing :80-101,
_events_to_analysis_dimensions :104-106, InterpolateAndExtrapolate1D :109-147, arrays_to_grid
:150-153); the implementations are independent.
"""
import copy
import hashlib
import os
import pickle as _pickle
import tempfile

import numpy as np

try:                                    # the reference serialises with dill; plain pickle otherwise
    import dill as _serializer
except ImportError:                     # pragma: no cover
    _serializer = _pickle

__all__ = ['inherit_docstring_from', 'combine_dicts', 'data_file_name', 'find_file_in_folders',
           'read_pickle', 'save_pickle', 'hashablize', 'deterministic_hash',
           'InterpolateAndExtrapolate1D', 'arrays_to_grid']


def inherit_docstring_from(cls):
    """Decorator: copy the docstring of the same-named attribute of `cls`."""
    def decorate(fn):
        fn.__doc__ = getattr(cls, fn.__name__).__doc__
        return fn
    return decorate


def combine_dicts(*dicts, exclude=(), deep_copy=False):
    """Merge dicts left to right (later wins), drop `exclude` keys, optionally deep-copying inputs."""
    merged = {}
    for d in dicts:
        merged.update(copy.deepcopy(d) if deep_copy else d)
    for key in exclude:
        merged.pop(key, None)
    return merged


def find_file_in_folders(filename, folders):
    """First existing folder/filename (no recursion); FileNotFoundError otherwise."""
    for folder in ([folders] if isinstance(folders, str) else folders):
        candidate = os.path.join(folder, filename)
        if os.path.exists(candidate):
            return candidate
    raise FileNotFoundError(filename)


def data_file_name(filename, data_dirs=None):
    if os.path.exists(filename):
        return filename
    if data_dirs is not None:
        return find_file_in_folders(filename, data_dirs)
    return FileNotFoundError(filename)     # (sic) the reference returns the exception object here


def read_pickle(filename):
    with open(filename, 'rb') as f:
        return _serializer.load(f)


def save_pickle(stuff, filename):
    """Write atomically: dump to a temporary file in the target directory, then rename."""
    directory = os.path.dirname(filename)
    if directory:
        os.makedirs(directory, exist_ok=True)
    fd, tmp = tempfile.mkstemp(dir=directory or '.')
    try:
        with os.fdopen(fd, 'wb') as f:
            _serializer.dump(stuff, f)
        os.replace(tmp, filename)
    except BaseException:
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise


def hashablize(obj):
    """Recursively turn dicts / arrays / iterables into tuples so that the result is hashable."""
    try:
        hash(obj)
        return obj
    except TypeError:
        pass
    if isinstance(obj, dict):
        return tuple((k, hashablize(v)) for k, v in sorted(obj.items()))
    if isinstance(obj, np.ndarray):
        return tuple(obj.tolist())
    if hasattr(obj, '__iter__'):
        return tuple(hashablize(x) for x in obj)
    raise TypeError("Can't hashablize object of type %r" % type(obj))


def deterministic_hash(thing):
    """sha1 of the pickled hashablized object -- same digest as the reference for the same config."""
    return hashlib.sha1(_pickle.dumps(hashablize(thing))).hexdigest()


def _events_to_analysis_dimensions(events, analysis_space):
    return [events[name] for name, _ in analysis_space]


class InterpolateAndExtrapolate1D(object):
    """Piecewise-linear 1-D interpolation that holds the end values outside the data range."""

    def __init__(self, points, values):
        points = np.atleast_1d(np.asarray(points, dtype=float))
        values = np.atleast_1d(np.asarray(values, dtype=float))
        assert len(points) == len(values)
        self.points, self.values = points, values
        self.min, self.max = points.min(), points.max()

    def __call__(self, x):
        scalar = np.ndim(x) == 0
        xs = np.clip(np.atleast_1d(np.asarray(x, dtype=float)), self.min, self.max)
        if len(self.points) == 1:
            out = np.full(len(xs), self.values[0])
        else:
            order = np.argsort(self.points)
            out = np.interp(xs, self.points[order], self.values[order])
        return out[0] if scalar else out


def arrays_to_grid(arrs):
    """n 1-d arrays -> (n+1)-d array, 'ij' indexing, last axis = the coordinates of each grid point."""
    return np.stack(np.meshgrid(*arrs, indexing='ij'), axis=-1)
