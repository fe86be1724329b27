#!/usr/bin/env python
"""Throughput of the drop-in call ViT_opencl(ImageData[], Network[], float**) with the reference's
data layout (one malloc'd buffer per image, pageable): python tools/dropin_throughput.py [n] [precision]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
os.environ["VITB200_PRECISION"] = sys.argv[2] if len(sys.argv) > 2 else "bf16"
blobs = pkg.synth.model_blobs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "Network"))
base = pkg.synth.synthetic_images(64, 224, seed=3)
images = np.ascontiguousarray(base[np.arange(n) % 64])
for rep in range(3):
    t0 = time.perf_counter()
    probs = pkg.vit_opencl(images, blobs)
    dt = time.perf_counter() - t0
    print(f"call {rep}: {n} images in {dt:.3f} s = {n / dt:.0f} images/s (includes device bring-up + weight upload)")
