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


# -- pdf-cache format (blueice/source.py:106-128,155-160; blueice/utils.py:65-77) -----------------------------------------
# A reference cache file is a (dill) pickle of {attribute: value} whose histograms are `multihist.Histdd` instances.
# Reading: such pickles load here even without multihist -- its histogram classes are mapped onto blueice_b200.hist.Histdd,
# which takes over their instance dict (histogram, bin_edges, dimensions, axis_names).  Writing: when multihist can be
# imported, histograms are written as multihist objects, so the reference can load caches made here.  (The state of real
# multihist objects is restated from memory of multihist 0.6.x -- SURVEY.md 8c; pinned against the test stand-in only.)
_MULTIHIST_CLASSES = ('Histdd', 'Hist1d', 'MultiHistBase')


class _CacheUnpickler(getattr(_serializer, 'Unpickler', _pickle.Unpickler)):
    def find_class(self, module, name):
        if module.split('.')[0] == 'multihist' and name in _MULTIHIST_CLASSES:
            try:
                return super().find_class(module, name)
            except (ImportError, AttributeError):
                from .hist import Histdd
                return Histdd
        return super().find_class(module, name)


def _own_histograms(obj):
    """multihist histograms anywhere in a cache dict -> blueice_b200.hist.Histdd (the device paths expect those)."""
    from .hist import Histdd
    if isinstance(obj, dict):
        return {k: _own_histograms(v) for k, v in obj.items()}
    if type(obj).__module__.split('.')[0] == 'multihist' and hasattr(obj, 'histogram') and hasattr(obj, 'bin_edges'):
        edges = obj.bin_edges if getattr(obj, 'dimensions', 2) != 1 or isinstance(obj.bin_edges, (list, tuple)) \
            else [obj.bin_edges]
        return Histdd.from_histogram(obj.histogram, edges, axis_names=getattr(obj, 'axis_names', None))
    return obj


def _reference_histograms(obj):
    """blueice_b200.hist.Histdd anywhere in a cache dict -> multihist.Histdd when multihist is importable."""
    from .hist import Histdd
    if isinstance(obj, dict):
        return {k: _reference_histograms(v) for k, v in obj.items()}
    if type(obj) is Histdd:
        try:
            import multihist
        except ImportError:
            return obj
        out = multihist.Histdd(bins=obj.bin_edges, axis_names=obj.axis_names)
        out.histogram = np.array(obj.histogram, dtype=float)
        return out
    return obj


def read_pickle(filename):
    with open(filename, 'rb') as f:
        return _own_histograms(_CacheUnpickler(f).load())


def save_pickle(stuff, filename):
    """Write atomically: dump to a temporary file in the target directory, then rename."""
    directory = os.path.dirname(filename)
    if directory:
        os.makedirs(directory, exist_ok=True)
    stuff = _reference_histograms(stuff)
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
