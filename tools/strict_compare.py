#!/usr/bin/env python
"""Compare two Main.c-style result files ("[i] label: L / prob: P"), any image count.

    python tools/strict_compare.py Data/opencl_result.txt Data/answer_result.txt [--tol 1e-6]

Exit status = number of errors (capped at 255), like the value comparator() hands back to Main.c
(R/Main.c:75-82).  The default tolerance is the reference's 0.01; --tol tightens it.
"""
import argparse
import importlib.util
import os
import sys

here = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("vit_compare", os.path.join(here, "..", "vit-with-opencl_b200", "compare.py"))
cmp = importlib.util.module_from_spec(spec)
sys.modules["vit_compare"] = cmp
spec.loader.exec_module(cmp)

ap = argparse.ArgumentParser()
ap.add_argument("result")
ap.add_argument("answer")
ap.add_argument("--tol", type=float, default=0.01)
ap.add_argument("--count", type=int, default=None)
a = ap.parse_args()
rep = cmp.compare_result_files(open(a.result).read(), open(a.answer).read(), a.tol, a.count)
for m in rep.messages:
    print(m, file=sys.stderr)
print(f"{rep.stats.get('lines', 0)} lines, {rep.errors} errors, max |dprob| {rep.stats.get('max_abs_dprob', 0):.6f}")
sys.exit(min(rep.errors, 255))
