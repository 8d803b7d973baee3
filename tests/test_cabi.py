"""The C-ABI library loads on a CPU-only box and exports exactly what include/blueice_b200.h declares.
No compute calls: argument validation happens before any CUDA call, so error paths are testable here."""
import ctypes
import os
import re

import numpy as np

from blueice_b200 import _build, _cabi

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "blueice_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bi_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    path = _build.build()
    assert os.path.dirname(path) == os.path.dirname(_build.__file__)
    assert os.path.exists(path)


def test_every_declared_symbol_is_exported_and_bound():
    lib = _cabi.load()
    names = declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), "symbol %s declared in the header is not exported" % name
        assert name in _cabi.SIGNATURES, "symbol %s has no ctypes signature" % name
    for name in _cabi.SIGNATURES:
        assert name in names, "ctypes signature %s is not declared in the header" % name


def test_header_constants_match_python_mirror():
    text = open(HEADER).read()
    consts = dict(re.findall(r"#define\s+BI_([A-Z0-9_]+)\s+(-?\d+)", text))
    assert int(consts["MAX_DIMS"]) == _cabi.MAX_DIMS
    assert int(consts["MAX_SOURCES"]) == _cabi.MAX_SOURCES
    assert int(consts["EVENT_BLOCK"]) == _cabi.EVENT_BLOCK
    assert int(consts["SUPERBLOCK"]) == _cabi.SUPERBLOCK
    assert int(consts["MAX_AXIS_POINTS"]) == _cabi.MAX_AXIS_POINTS
    assert int(consts["ABI_VERSION"]) == _cabi.load().bi_abi_version()


def test_superblock_count():
    lib = _cabi.load()
    assert lib.bi_num_superblocks(0) == 0
    assert lib.bi_num_superblocks(1) == 1
    assert lib.bi_num_superblocks(512) == 1
    assert lib.bi_num_superblocks(513) == 2
    assert lib.bi_num_superblocks(100000) == 196
    # K4 scratch: sum_t [P] + point constants [PC, 4] + two [PC, 16 * n_super] block-sum arrays + the per-bin scratch (stored
    # terms [P, 8, tb_ld] for a few points, t_b [PC, tb_ld] otherwise) + the schedule
    n_blocks, tb_ld = 32, 1024
    assert lib.bi_binned_scratch_doubles(3, 1000) == 3 + 4 * 3 + 2 * 3 * n_blocks + 3 * 8 * tb_ld + (5 * 3 + 12 + 1) // 2 + 8
    assert lib.bi_binned_scratch_doubles(0, 1000) == 0
    assert lib.bi_binned_sum_t_offset(3, 1000) == 0
    assert lib.bi_unbinned_small_ok(1, 1, 1, 1012) == 1 and lib.bi_unbinned_small_ok(2, 2, 4096, 100000) == 0
    assert lib.bi_peer_exchange_words(8, 11) == 2 * 8 * 11 + 8 + 4


def test_argument_validation_reports_errors_without_a_gpu():
    lib = _cabi.load()
    n_anchors = _cabi.as_i32([3] * 7)
    axes = _cabi.as_f64(np.zeros(21))
    rc = lib.bi_point_setup(7, _cabi.host_ptr(n_anchors), _cabi.host_ptr(axes), 1, 1, None, None, None, None, None,
                            None, None, None, None, None, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"n_dims" in lib.bi_last_error()
    # non-increasing anchor axis
    n_anchors = _cabi.as_i32([3])
    axes = _cabi.as_f64([0., 2., 1.])
    rc = lib.bi_point_setup(1, _cabi.host_ptr(n_anchors), _cabi.host_ptr(axes), 1, 1, None, None, None, None, None,
                            None, None, None, None, None, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"increasing" in lib.bi_last_error()
    # odd leading dimension of the anchor tensor
    rc = lib.bi_unbinned_partials_stream(ctypes.c_void_p(16), 7, 7, 1, 1, None, 1, None, None, None, None, 1e-12,
                                         None, None)
    assert rc == -1 and b"ld_events" in lib.bi_last_error()
    # unknown lookup method
    rc = lib.bi_hist_lookup(None, 1, 1, _cabi.host_ptr(_cabi.as_i32([4])), _cabi.host_ptr(_cabi.as_f64(np.arange(5.))),
                            None, 0, 0, 9, None, 0, None, None)
    assert rc == -1 and b"method" in lib.bi_last_error()
    # a long contraction without the chunk-major coefficient buffer, and more terms than any kernel takes
    p = ctypes.c_void_p(256)
    rc = lib.bi_unbinned_partials_mma(p, 64, 64, 160, 5, p, p, p, p, p, p, p, p, 1e-12, p, -1, None, None, 10, None, None)
    assert rc == -1 and b"coef_chunks_dev" in lib.bi_last_error()
    rc = lib.bi_unbinned_partials_mma(p, 64, 64, _cabi.MMA_MAX_TERMS + 1, 5, p, p, p, p, p, p, p, p, 1e-12, p, -1, None, None,
                                      10, p, None)
    assert rc == -1 and b"contraction terms" in lib.bi_last_error()


def test_kernel_choice_and_workspace_of_long_contractions():
    """Host-side rules of K2 (no device needed): contractions of more than 32 terms run the K-chunk kernel -- units of 64
    points, a chunk-major copy of the coefficients ([chunk][slot][36], 64 slots of padding) as the last workspace region."""
    lib = _cabi.load()
    for k in (1, 8, 16):
        assert lib.bi_mma_unit_points(k, 4096) == 32 and lib.bi_mma_coef_chunks_doubles(k, 100) == 0
    assert lib.bi_mma_unit_points(32, 4096) == 16 and lib.bi_mma_coef_chunks_doubles(32, 100) == 0
    for k in (33, 48, 128, 129, 160, _cabi.MMA_MAX_TERMS):
        assert lib.bi_mma_unit_points(k, 4096) == 64 and lib.bi_mma_unit_points(k, 1) == 64
        chunks = -(-k // 32)
        assert lib.bi_mma_coef_chunks_doubles(k, 100) == chunks * (100 + 64) * 36
    off = np.zeros(15, dtype=np.int64)
    assert lib.bi_unbinned_workspace_layout(5, 5, 160, 4096, 50000, _cabi.host_ptr(off)) == 0
    assert np.all(np.diff(off) >= 0) and np.all(off % 256 == 0)
    assert off[14] == lib.bi_unbinned_workspace_bytes(5, 5, 160, 4096, 50000)
    assert off[14] - off[13] >= 8 * lib.bi_mma_coef_chunks_doubles(160, 4096) > 0
    off8 = np.zeros(15, dtype=np.int64)
    assert lib.bi_unbinned_workspace_layout(2, 2, 8, 4096, 99957, _cabi.host_ptr(off8)) == 0
    assert off8[14] == off8[13]                                      # short contractions: no such region


def test_ctypes_signatures_have_the_header_s_parameter_counts_and_kinds():
    """Every prototype of include/blueice_b200.h against its ctypes mirror: number of parameters, and pointer / integer /
    floating-point kind of each (a mismatch corrupts the call silently)."""
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = dict(re.findall(r"\b(bi_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S))
    checked = 0
    for name, (restype, argtypes) in _cabi.SIGNATURES.items():
        assert name in protos, name
        params = [p.strip() for p in protos[name].split(",")] if protos[name].strip() not in ("", "void") else []
        assert len(params) == len(argtypes), "%s: header has %d parameters, ctypes %d" % (name, len(params), len(argtypes))
        for p, t in zip(params, argtypes):
            if "*" in p:
                assert t in (ctypes.c_void_p, ctypes.c_char_p), (name, p, t)
            elif re.search(r"\b(double|float)\b", p):
                assert t is ctypes.c_double, (name, p, t)
            elif re.search(r"\bint64_t\b|\buint64_t\b", p):
                assert ctypes.sizeof(t) == 8 and t is not ctypes.c_double and t is not ctypes.c_void_p, (name, p, t)
            else:
                assert ctypes.sizeof(t) == 4, (name, p, t)
        checked += 1
    assert checked >= 40
