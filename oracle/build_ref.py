"""Recipe for oracle/_ref: the UNMODIFIED reference package, copied where it lies under /root/reference, next to the two
test-only stand-ins for its uninstallable dependencies (multihist, atomicwrites; tests/golden/_shims).

    python oracle/build_ref.py          # or __graft_entry__.build()

TEST INFRASTRUCTURE ONLY: oracle/_ref is git-ignored (a build product, never committed), not gpurun-ignored (it travels
to the GPU box with the snapshot, where /root/reference does not exist).  bench.py's `--impl reference` arm and its
cpu_baseline leg time it (cpu_baseline.kind = "reference"); the tests validate the oracle restatement against golden
vectors produced from it (tests/golden/make_golden.py).  Nothing under blueice_b200/ ever imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/blueice"
SHIMS = os.path.join(os.path.dirname(HERE), "tests", "golden", "_shims")
DEST = os.path.join(HERE, "_ref")


def build_ref(force=False):
    """Copy the reference package and the shims to oracle/_ref.  Returns the path, or None when /root/reference is absent
    (the GPU box: the prebuilt copy that travelled with the snapshot is used as it is)."""
    if not os.path.isdir(REF_SRC):
        return DEST if os.path.isdir(os.path.join(DEST, "blueice")) else None
    marker = os.path.join(DEST, "blueice", "__init__.py")
    if os.path.exists(marker) and not force:
        return DEST
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    shutil.copytree(REF_SRC, os.path.join(DEST, "blueice"), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for name in os.listdir(SHIMS):
        src = os.path.join(SHIMS, name)
        if os.path.isdir(src):
            shutil.copytree(src, os.path.join(DEST, name), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return DEST


def import_reference():
    """Import the reference package from oracle/_ref (ahead of everything else on sys.path).  Returns the module, or None."""
    if not os.path.isdir(os.path.join(DEST, "blueice")):
        return None
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import blueice
    return blueice


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
