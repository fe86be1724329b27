#!/usr/bin/env python
"""Batch-1 resident latency, p50 / p10 over N forwards (graph replay): python tools/b1_latency.py [fp32|bf16] [iters]
A/B switches are environment variables read by the library (VITB200_FP32_SPLITK=0, VITCU_ATTN_SMALL=0, VITCU_L2_PREFETCH=0)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "fp32"
prec = pkg.FP32 if name == "fp32" else pkg.BF16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
blobs = pkg.synth.model_blobs(None, 224, seed=7)
x = pkg.synth.synthetic_images(1, 224, seed=1)
with pkg.Engine(0, 224, prec, max_batch=1) as e:
    e.load_weights(blobs)
    e.stage(x)
    for _ in range(20):
        e.forward_resident(1)
    ms = sorted(e.forward_resident(1) for _ in range(iters))
    print(f"{name} batch-1 resident: p50 {ms[len(ms) // 2]:.4f} ms  p10 {ms[len(ms) // 10]:.4f}  min {ms[0]:.4f}  "
          f"launches {e.kernels_per_forward}  env "
          + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith(("VITB200_", "VITCU_"))))
