"""ctypes binding of the C-ABI library (include/blueice_b200.h).

There is no CPU fallback: if the shared library is missing or the CUDA device is unavailable the
product path raises.  Pointers are passed as integers (torch `tensor.data_ptr()`), host descriptor
arrays as NumPy arrays.
"""
import ctypes
import os

import numpy as np

from . import _build

_c_void_p = ctypes.c_void_p
_i32 = ctypes.c_int32
_i64 = ctypes.c_int64
_f64 = ctypes.c_double

# name -> (restype, argtypes); the order follows include/blueice_b200.h
SIGNATURES = {
    "bi_last_error": (ctypes.c_char_p, []),
    "bi_abi_version": (ctypes.c_int, []),
    "bi_num_superblocks": (_i64, [_i64]),
    "bi_point_setup": (ctypes.c_int, [_i32, _c_void_p, _c_void_p, _i32, _i64,
                                      _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                      _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                      _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "bi_unbinned_partials_stream": (ctypes.c_int, [_c_void_p, _i64, _i64, _i32, _i32, _c_void_p, _i64,
                                                   _c_void_p, _c_void_p, _c_void_p, _c_void_p, _f64,
                                                   _c_void_p, _c_void_p]),
    "bi_plan_max_cells": (_i64, []),
    "bi_unbinned_plan": (ctypes.c_int, [_i32, _c_void_p, _i64, _c_void_p, _c_void_p, _i32, _i64, _i32, _i32,
                                        _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "bi_unbinned_partials_mma": (ctypes.c_int, [_c_void_p, _i64, _i64, _i32, _i32, _c_void_p, _c_void_p,
                                                _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                                _f64, _c_void_p, _i32, _c_void_p, _c_void_p, _i64, _c_void_p,
                                                _c_void_p]),
    "bi_mma_unit_points": (_i32, [_i32, _i64]),
    "bi_mma_coef_chunks_doubles": (_i64, [_i32, _i64]),
    "bi_unbinned_workspace_bytes": (_i64, [_i32, _i32, _i32, _i64, _i64]),
    "bi_unbinned_workspace_layout": (ctypes.c_int, [_i32, _i32, _i32, _i64, _i64, _c_void_p]),
    "bi_unbinned_ll_batch": (ctypes.c_int, [_i32, _c_void_p, _c_void_p, _i32, _i64,
                                            _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                            _c_void_p, _i64, _i64, _f64, _i32, _c_void_p, _i64,
                                            _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "bi_unbinned_small_ok": (_i32, [_i32, _i32, _i64, _i64]),
    "bi_unbinned_ll_small": (ctypes.c_int, [_i32, _c_void_p, _c_void_p, _i32, _i64,
                                            _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                            _c_void_p, _i64, _i64, _f64,
                                            _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "bi_sourcewise_terms": (_i32, [_i32, _c_void_p]),
    "bi_point_setup_sourcewise": (ctypes.c_int, [_i32, _c_void_p, _c_void_p, _i32, _c_void_p, _c_void_p, _i64,
                                                 _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                                 _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                                 _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "bi_unbinned_ll_batch_sourcewise": (ctypes.c_int, [_i32, _c_void_p, _c_void_p, _i32, _c_void_p, _c_void_p, _i64,
                                                       _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                                       _c_void_p, _c_void_p, _i64, _i64, _f64, _i32, _c_void_p, _i64,
                                                       _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "bi_unbinned_ps_terms": (ctypes.c_int, [_c_void_p, _i64, _i64, _i32, _i32, _c_void_p, _c_void_p, _c_void_p,
                                            _c_void_p, _c_void_p, _i64, _c_void_p]),
    "bi_unbinned_finalize": (ctypes.c_int, [_c_void_p, _i64, _c_void_p, _c_void_p, _i64, _c_void_p, _c_void_p,
                                            _c_void_p]),
    "bi_unbinned_ps": (ctypes.c_int, [_c_void_p, _i64, _i64, _i32, _i32, _c_void_p, _c_void_p, _c_void_p,
                                      _i64, _c_void_p]),
    "bi_hist_lookup": (ctypes.c_int, [_c_void_p, _i64, _i32, _c_void_p, _c_void_p, _c_void_p, _i64, _i64,
                                      _i32, _c_void_p, _i64, _c_void_p, _c_void_p]),
    "bi_template_prepare_events": (ctypes.c_int, [_i32, _c_void_p, _c_void_p, _i32, _c_void_p, _i64, _i64,
                                                  _c_void_p, _c_void_p, _i64, _c_void_p]),
    "bi_template_partials": (ctypes.c_int, [_c_void_p, _i64, _i64, _i32, _c_void_p, _i32, _c_void_p, _c_void_p, _i64,
                                            _c_void_p, _i32, _i32, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                            _c_void_p, _c_void_p, _i64, _i32, _c_void_p, _c_void_p, _c_void_p, _i64,
                                            _c_void_p, _c_void_p, _f64, _c_void_p, _c_void_p]),
    "bi_template_finalize": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _i64, _i64,
                                            _c_void_p, _c_void_p, _c_void_p]),
    "bi_template_mix": (ctypes.c_int, [_c_void_p, _i64, _i64, _i32, _c_void_p, _i32, _i32, _c_void_p, _c_void_p,
                                       _c_void_p, _c_void_p, _i64, _c_void_p, _c_void_p]),
    "bi_mixture_partials": (ctypes.c_int, [_c_void_p, _i32, _c_void_p, _i32, _c_void_p, _c_void_p, _i64, _c_void_p,
                                           _c_void_p, _i64, _i32, _c_void_p, _c_void_p, _c_void_p, _i64, _c_void_p,
                                           _c_void_p, _f64, _c_void_p, _c_void_p]),
    "bi_template_workspace_bytes": (_i64, [_i32, _i32, _i64, _i64, _i64, _i64, _i32]),
    "bi_template_ll_batch": (ctypes.c_int, [_i32, _c_void_p, _c_void_p, _i32, _i64,
                                            _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                            _c_void_p, _i64, _i64, _i32, _c_void_p, _i32, _i32,
                                            _c_void_p, _c_void_p, _i64, _c_void_p,
                                            _i64, _i32, _c_void_p, _c_void_p, _c_void_p, _i64, _c_void_p, _c_void_p,
                                            _i64, _i64, _i64, _f64, _c_void_p, _i64,
                                            _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "bi_template_bm_supported": (ctypes.c_int, [_i32, _i32, _i32, _c_void_p, _i32, _i64]),
    "bi_template_bm_record_doubles": (_i64, [_i32, _i32]),
    "bi_template_bm_chunk": (_i32, []),
    "bi_template_bm_density": (ctypes.c_int, [_c_void_p, _i64, _i32, _i32, _c_void_p, _i32, _i64,
                                              _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                              _c_void_p, _c_void_p, _c_void_p, _i64,
                                              _c_void_p, _c_void_p, _c_void_p, _i64, _c_void_p, _c_void_p, _c_void_p]),
    "bi_template_ll_toys_bm": (ctypes.c_int, [_i32, _c_void_p, _c_void_p, _i32, _i64,
                                              _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                              _c_void_p, _i64, _i64, _c_void_p, _i64, _i32, _c_void_p,
                                              _c_void_p, _c_void_p, _i64, _c_void_p,
                                              _c_void_p, _c_void_p, _c_void_p, _i64,
                                              _c_void_p, _c_void_p, _c_void_p, _i64,
                                              _i64, _c_void_p, _c_void_p, _c_void_p, _i64, _c_void_p, _c_void_p,
                                              _i64, _i64, _f64, _c_void_p, _i64, _c_void_p, _c_void_p,
                                              _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "bi_toy_counts": (ctypes.c_int, [_i32, _i64, _i64, _c_void_p, _i32, ctypes.c_uint64, _c_void_p, _c_void_p]),
    "bi_toy_events": (ctypes.c_int, [_i32, _c_void_p, _c_void_p, _i32, _c_void_p, _i64, _i64, _c_void_p, _c_void_p,
                                     _i64, ctypes.c_uint64, _c_void_p, _i64, _c_void_p, _c_void_p]),
    "bi_histogramdd": (ctypes.c_int, [_i32, _c_void_p, _c_void_p, _c_void_p, _i64, _i64, _c_void_p,
                                      _c_void_p, _c_void_p]),
    "bi_binned_scratch_doubles": (_i64, [_i64, _i64]),
    "bi_binned_sum_t_offset": (_i64, [_i64, _i64]),
    "bi_binned_ll_batch": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _i64, _i64, _i32, _i32, _i32,
                                          _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                          _i64, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "bi_histogramdd_toys": (ctypes.c_int, [_i32, _c_void_p, _c_void_p, _c_void_p, _i64, _i64, _c_void_p, _i64,
                                           _c_void_p, _i64, _c_void_p]),
    "bi_binned_ll_batch_toys": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _i64, _i64, _i32, _i32, _i32,
                                               _c_void_p, _c_void_p, _i64, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                               _i64, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "bi_binned_pmfs": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _i64, _i64, _i32, _i32, _i32,
                                      _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                      _i64, _c_void_p]),
    "bi_peer_exchange_words": (_i64, [_i32, _i64]),
    "bi_peer_exchange": (ctypes.c_int, [_c_void_p, _i64, _i64, _c_void_p, _i32, _i32, _i32, _c_void_p, _c_void_p,
                                        _c_void_p, _c_void_p]),
    "bi_peer_gather_status": (ctypes.c_int, [_c_void_p, _i64, _i64, _c_void_p, _i32, _i32, _c_void_p, _c_void_p, _c_void_p,
                                             _c_void_p]),
    "bi_peer_broadcast": (ctypes.c_int, [_c_void_p, _i64, _c_void_p, _i32, _i64, _c_void_p]),
    "bi_bench_fp64_fma": (ctypes.c_int, [_i64, _i32, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "bi_bench_fp64_mma": (ctypes.c_int, [_i64, _i32, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "bi_bench_stream_read": (ctypes.c_int, [_c_void_p, _i64, _c_void_p, _c_void_p, _c_void_p]),
}

# mirror of the #defines in include/blueice_b200.h
MAX_DIMS = 6
MAX_SOURCES = 64
MAX_SPACE_DIMS = 4
MAX_AXIS_POINTS = 256
EVENT_BLOCK = 32
SUPERBLOCK = 512
STREAM_MAX_CORNERS = 32
MMA_MAX_TERMS = 4096
PLAN_MAX_CELLS = 16384
TS_MAX_TERMS = 256
TS_GROUP_POINTS = 8
MIX_GROUP_POINTS = 8
MIX_GROUP_POINTS_WIDE = 16
POINT_OUT_OF_RANGE = 1
POINT_UNPHYSICAL = 2
LOOKUP_LINEAR = 0
LOOKUP_PIECEWISE = 1
BB_ROOT1_POSITIVE = 1
BB_NEGATIVE_A = 2


class BlueiceB200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""


_lib = None


def library_path():
    return os.environ.get("BLUEICE_B200_LIB") or _build.LIB_PATH     # override: kernel-variant experiments


def load():
    """Load (once) and return the ctypes library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            "blueice_b200: %s is missing. Build it with `python -m blueice_b200._build` "
            "(needs nvcc; sm_100a). There is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)             # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.bi_abi_version() != 1:
        raise ImportError("blueice_b200: ABI version mismatch, rebuild the library")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().bi_last_error()
        raise BlueiceB200Error("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))


def host_ptr(arr):
    """Pointer to a C-contiguous NumPy array kept alive by the caller."""
    if arr is None:
        return None
    assert arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(ctypes.c_void_p)


def dev_ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def as_i32(values):
    return np.ascontiguousarray(np.asarray(values, dtype=np.int32))


def as_f64(values):
    return np.ascontiguousarray(np.asarray(values, dtype=np.float64))
