"""Per-kernel SASS opcode histogram of the built library (evidence that the hot kernels are the hand-written sm_100a ones:
DMMA = FP64 tensor instruction, UTMALDG = tiled TMA load, UBLKCP = 1-D TMA bulk copy, SYNCS = mbarrier operations).

    python profiles/sass_histogram.py > profiles/r2_sass_opcodes.md        (needs cuobjdump; runs without a GPU)
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "blueice_b200", "libblueice_b200.so")
WATCH = ["DMMA", "DFMA", "DMUL", "DADD", "MUFU", "UTMALDG", "UBLKCP", "SYNCS", "LDG", "LDS", "STG", "STS", "SHFL", "BAR",
         "ATOM", "RED", "NOP"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        return out if len(out) == len(names) else names
    except OSError:
        return names


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    names = list(kernels)
    pretty = demangle(names)
    print("# SASS opcode histogram per kernel (round 2)\n")
    print("`cuobjdump -sass blueice_b200/libblueice_b200.so`, opcode = mnemonic before the first dot; static instruction "
          "counts per kernel (all template instantiations listed).\n")
    print("| kernel | total | " + " | ".join(WATCH) + " |")
    print("|---|---|" + "---|" * len(WATCH))
    totals = collections.Counter()
    for name, nice in zip(names, pretty):
        c = kernels[name]
        nice = re.sub(r"\(.*", "", nice)
        print("| `%s` | %d | %s |" % (nice[:90], c["_total"], " | ".join(str(c.get(w, 0)) for w in WATCH)))
        totals.update(c)
    print("| **all kernels** | %d | %s |" % (totals["_total"], " | ".join(str(totals.get(w, 0)) for w in WATCH)))
    print("\nNo `UTC*MMA` / `TCGEN05` / TMEM instruction appears: the path is float64 throughout (DESIGN.md section 5), "
          "and tcgen05 has no FP64 type; `HMMA`/`IMMA` count: %d." % (totals.get("HMMA", 0) + totals.get("IMMA", 0)))


if __name__ == "__main__":
    sys.exit(main())
