"""One-off device operations with host inputs/outputs (thin wrappers over the C-ABI).

Used by Source.pdf and BinnedLogLikelihood.set_data; the batched likelihood path lives in engine.py.
No CPU fallback: without a CUDA device these raise.
"""
import ctypes

import numpy as np

from . import _cabi
from .engine import require_cuda, round_up

_METHODS = {'linear': _cabi.LOOKUP_LINEAR, 'piecewise': _cabi.LOOKUP_PIECEWISE}


def _coords_to_device(torch, coordinate_arrays, device):
    coords = np.ascontiguousarray(np.asarray([np.asarray(c, dtype=np.float64).reshape(-1)
                                              for c in coordinate_arrays], dtype=np.float64))
    return torch.from_numpy(coords).to(device), coords


def hist_lookup(templates, edges_list, coordinate_arrays, method='linear', return_bin_index=False, device=None):
    """K3 on host arrays.  templates [T, *bins] -> values [T, N] (bi_hist_lookup; source.py:219-246)."""
    torch = require_cuda()
    lib = _cabi.load()
    device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    templates = np.ascontiguousarray(np.asarray(templates, dtype=np.float64))
    T = templates.shape[0]
    n_space = len(edges_list)
    if len(coordinate_arrays) != n_space:
        raise ValueError("The requested sample points have dimension %d but the histogram has dimension %d"
                         % (len(coordinate_arrays), n_space))
    coords_d, coords = _coords_to_device(torch, coordinate_arrays, device)
    n = coords.shape[1]
    if method == 'linear' and n and np.isnan(coords).any():
        # scipy's RegularGridInterpolator bounds check rejects NaN (source.py:240)
        raise ValueError("One of the requested xi is out of bounds in dimension 0")
    n_bins = _cabi.as_i32([len(e) - 1 for e in edges_list])
    edges = _cabi.as_f64(np.concatenate([np.asarray(e, dtype=np.float64) for e in edges_list]))
    ld = max(round_up(n, 2), 2)
    out = torch.empty((T, ld), dtype=torch.float64, device=device)
    bin_index = torch.empty(max(n, 1), dtype=torch.int32, device=device) if return_bin_index else None
    tmpl = torch.from_numpy(templates.reshape(T, -1)).to(device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    rc = lib.bi_hist_lookup(_cabi.dev_ptr(tmpl), T, n_space, _cabi.host_ptr(n_bins), _cabi.host_ptr(edges),
                            _cabi.dev_ptr(coords_d), max(n, 1) if n == 0 else coords_d.shape[1], n,
                            _METHODS[method], _cabi.dev_ptr(out), ld, _cabi.dev_ptr(bin_index), stream)
    _cabi.check(rc, "bi_hist_lookup")
    values = out[:, :n].cpu().numpy()
    if return_bin_index:
        return values, bin_index[:n].cpu().numpy()
    return values


def histogramdd(edges_list, coordinate_arrays, return_bin_index=False, device=None):
    """np.histogramdd-compatible event binning on device (bi_histogramdd; likelihood.py:604-609)."""
    torch = require_cuda()
    lib = _cabi.load()
    device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    n_space = len(edges_list)
    n_bins = _cabi.as_i32([len(e) - 1 for e in edges_list])
    edges = _cabi.as_f64(np.concatenate([np.asarray(e, dtype=np.float64) for e in edges_list]))
    coords_d, coords = _coords_to_device(torch, coordinate_arrays, device)
    n = coords.shape[1] if coords.ndim == 2 else 0
    counts = torch.zeros(int(np.prod(n_bins)), dtype=torch.int64, device=device)
    bin_index = torch.empty(max(n, 1), dtype=torch.int32, device=device) if return_bin_index else None
    if n:
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        rc = lib.bi_histogramdd(n_space, _cabi.host_ptr(n_bins), _cabi.host_ptr(edges), _cabi.dev_ptr(coords_d),
                                coords_d.shape[1], n, _cabi.dev_ptr(counts), _cabi.dev_ptr(bin_index), stream)
        _cabi.check(rc, "bi_histogramdd")
    hist = counts.cpu().numpy().astype(np.float64).reshape([int(b) for b in n_bins])
    if return_bin_index:
        return hist, bin_index[:n].cpu().numpy()
    return hist
