#!/usr/bin/env python
"""ms per 256-image step as a function of how long the GPU has been kept busy (power/clock effects)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
blobs = pkg.synth.model_blobs(None, 224, seed=7)
x = pkg.synth.synthetic_images(256, 224, seed=1)
with pkg.Engine(0, 224, pkg.BF16, max_batch=256) as e:
    e.load_weights(blobs)
    e.stage(x)
    for _ in range(3):
        e.forward_resident(256)
    for steps in (1, 2, 5, 10, 20, 50, 100, 1, 2, 5):
        time.sleep(1.0)  # let the chip cool / clocks recover
        ms = e.time_resident(256, steps) / steps
        print(f"{steps:4d} back-to-back steps after 1 s idle: {ms:.3f} ms/step")
