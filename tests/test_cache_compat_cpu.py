"""pdf-cache format compatibility (blueice/source.py:106-128,155-160), on the CPU:
a cache file written by the UNMODIFIED reference is found (same sha1 file name) and loaded by blueice_b200 without
multihist installed, and a cache file written by blueice_b200 (with multihist importable) is loaded by the reference.
The reference runs in a subprocess from oracle/_ref (or /root/reference) with the test stand-ins for multihist /
atomicwrites: the histogram state is therefore pinned against the stand-in, not against real multihist."""
import os
import pickle
import subprocess
import sys
import textwrap

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(REPO, "tests", "golden", "_shims")


def _reference_root():
    for root in (os.path.join(REPO, "oracle", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(root, "blueice")):
            return root
    return None


def _run(code, cwd, path):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(path))
    res = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], cwd=cwd, env=env, capture_output=True, text=True,
                         timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    return res.stdout


CONFIG = """
import numpy as np
np.random.seed(5)
from {pkg}.test_helpers import conf_for_test
from {pkg}.model import Model
conf = conf_for_test(n_sources=1, mc=True, n_events_for_pdf=20000, mu=0.5)
m = Model(conf)
s = m.sources[0]
print("HASH", s.hash, s.from_cache, float(s.events_per_day), float(s.fraction_in_range), float(s._pdf_histogram.histogram.sum()))
"""


def test_reference_cache_is_loaded_and_written_compatibly(tmp_path):
    root = _reference_root()
    if root is None:
        pytest.skip("no reference package in this environment")
    # 1. the reference computes the source and writes ./pdf_cache/<sha1>
    out = _run(CONFIG.format(pkg="blueice"), str(tmp_path), [root, SHIMS])
    ref_hash, ref_from_cache, ref_rate, ref_frac, ref_sum = out.split("HASH")[1].split()
    assert ref_from_cache == "False"
    cache_file = tmp_path / "pdf_cache" / ref_hash
    assert cache_file.exists()
    # 2. blueice_b200 (no multihist on the path) finds the same file name and loads it
    out = _run(CONFIG.format(pkg="blueice_b200") + """
from blueice_b200.hist import Histdd
assert type(s._pdf_histogram) is Histdd and type(s._n_events_histogram) is Histdd
assert s.pdf_has_been_computed
np.save("loaded_hist.npy", s._pdf_histogram.histogram)
np.save("loaded_edges.npy", s._pdf_histogram.bin_edges[0])
""", str(tmp_path), [REPO])
    own_hash, own_from_cache, own_rate, own_frac, own_sum = out.split("HASH")[1].split()
    assert own_hash == ref_hash and own_from_cache == "True"
    assert (own_rate, own_frac, own_sum) == (ref_rate, ref_frac, ref_sum)
    # the stored histogram, read back with the stand-in classes, is what blueice_b200 loaded
    sys.path.insert(0, SHIMS)
    try:
        with open(cache_file, "rb") as f:
            stored = pickle.load(f)
    finally:
        sys.path.remove(SHIMS)
        for name in [k for k in sys.modules if k.split(".")[0] in ("multihist", "atomicwrites")]:
            del sys.modules[name]
    assert np.array_equal(np.load(tmp_path / "loaded_hist.npy"), stored["_pdf_histogram"].histogram)
    assert np.array_equal(np.load(tmp_path / "loaded_edges.npy"), stored["_pdf_histogram"].bin_edges[0])
    # 3. the other direction: blueice_b200 writes (multihist importable -> multihist objects), the reference loads
    other = tmp_path / "other"
    other.mkdir()
    out = _run(CONFIG.format(pkg="blueice_b200"), str(other), [REPO, SHIMS])
    w_hash, w_from_cache = out.split("HASH")[1].split()[:2]
    assert w_hash == ref_hash and w_from_cache == "False"
    out = _run(CONFIG.format(pkg="blueice") + """
import multihist
assert type(s._pdf_histogram) is multihist.Histdd
""", str(other), [root, SHIMS])
    assert out.split("HASH")[1].split()[1] == "True"
