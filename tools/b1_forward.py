#!/usr/bin/env python
"""Batch-1 forward (for ncu launch lists): python tools/b1_forward.py [fp32|bf16] [img] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
prec = pkg.FP32 if (len(sys.argv) < 2 or sys.argv[1] == "fp32") else pkg.BF16
img = int(sys.argv[2]) if len(sys.argv) > 2 else 224
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
blobs = pkg.synth.model_blobs(None, img, seed=7)
x = pkg.synth.synthetic_images(1, img, seed=1)
with pkg.Engine(0, img, prec, max_batch=1) as e:
    e.load_weights(blobs)
    e.stage(x)
    ms = [e.forward_resident(1) for _ in range(iters)]
print("ms per forward:", ms)
