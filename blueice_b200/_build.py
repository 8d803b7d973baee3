"""Build recipe for the C-ABI CUDA library (in-tree, sm_100a only).

    python -m blueice_b200._build          # or __graft_entry__.build()

Produces blueice_b200/libblueice_b200.so with nvcc.  `-fmad=false`: every fused multiply-add in the
kernels is an explicit fma(); everything else is separately rounded, which is what makes index,
weight and per-bin arithmetic bit-identical to the NumPy/SciPy reference.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_NAME = "libblueice_b200.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)
SOURCES = ["bi_util.cu", "bi_peer.cu", "bi_setup.cu", "bi_small.cu", "bi_unbinned.cu", "bi_unbinned_mma.cu", "bi_unbinned_wide.cu", "bi_plan.cu", "bi_lookup.cu", "bi_template.cu", "bi_template_bm.cu", "bi_toys.cu", "bi_binned.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def find_nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build %s" % LIB_NAME)


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    lib_time = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "blueice_b200.h"))
    return any(os.path.getmtime(d) > lib_time for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = find_nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    from concurrent.futures import ThreadPoolExecutor

    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    headers.append(os.path.join(HERE, "..", "include", "blueice_b200.h"))
    headers.append(os.path.abspath(__file__))
    header_time = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        if (not force and not verbose and os.path.exists(obj)
                and os.path.getmtime(obj) > max(header_time, os.path.getmtime(os.path.join(CSRC, src)))):
            return src, obj, None                                  # object is newer than its source and every header
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, res

    objects = []
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        for src, obj, res in pool.map(compile_one, SOURCES):
            if res is not None and (verbose or res.returncode != 0):
                sys.stderr.write(res.stdout + res.stderr)
            if res is not None and res.returncode != 0:
                raise RuntimeError("nvcc failed on %s" % src)
            objects.append(obj)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objects
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc link failed")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
