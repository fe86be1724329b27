#!/usr/bin/env python
"""Time the four per-layer launches of the BF16 tensor-core GEMM at M = batch*197
(CUDA events, default stream).  Usage: python tools/gemm_bench.py [batch] [reps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402
import bench  # noqa: E402

pkg = g.load_package()
L = pkg.lib()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
pkg.layer_check(L.vitcu_set_device(0))
per, flops, ms = bench.time_gemms(pkg, L, batch * 197, reps)
print(json.dumps({"M": batch * 197, "mode": os.environ.get("VITCU_GEMM_MODE", "pair"), "layer_ms": ms,
                  "tflops": flops / ms / 1e9, "per_launch": per}))
