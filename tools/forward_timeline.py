#!/usr/bin/env python
"""Where a forward's time goes, in situ: spans between consecutive launches of an eager forward (kernel + gap),
booked per kind, next to the captured-graph step time.  python tools/forward_timeline.py [batch] [bf16|fp32]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
L = pkg.lib()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = pkg.FP32 if (len(sys.argv) > 2 and sys.argv[2] == "fp32") else pkg.BF16
blobs = pkg.synth.model_blobs(None, 224, seed=7)
x = pkg.synth.synthetic_images(batch, 224, seed=1)
with pkg.Engine(0, 224, prec, max_batch=batch) as e:
    e.load_weights(blobs)
    e.stage(x)
    for _ in range(3):
        e.forward_resident(batch)
    graph_ms = e.time_resident(batch, 20) / 20
    ms, cnt = (C.c_float * 4)(), (C.c_int * 4)()
    acc = [0.0] * 4
    for _ in range(5):
        pkg._check(L.vitb200_profile_timeline(e.h, batch, ms, cnt))
        acc = [a + m for a, m in zip(acc, ms)]
    acc = [a / 5 for a in acc]
    gemm_ms, k = e.profile_gemms(batch, 5)
print(f"captured-graph step: {graph_ms:.3f} ms")
for name, a, c in zip(("GEMM", "attention", "LayerNorm", "other"), acc, cnt):
    print(f"  {name:10s} {c:3d} launches  {a:7.3f} ms  ({100 * a / sum(acc):4.1f} %)  {1e3 * a / max(c, 1):7.1f} us per launch incl. gap")
print(f"  eager forward with an event per launch: {sum(acc):.3f} ms;  GEMM kernels alone (event pair per launch): {gemm_ms:.3f} ms over {k} launches")
