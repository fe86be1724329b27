#!/usr/bin/env python
"""Where the drop-in call's time goes: engine bring-up, weight upload + packing, forward from the
reference's per-image (pageable) buffers vs from one pinned array."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
prec = pkg.BF16 if (len(sys.argv) < 3 or sys.argv[2] == "bf16") else pkg.FP32
blobs = pkg.synth.model_blobs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "Network"))
base = pkg.synth.synthetic_images(64, 224, seed=3)
images = np.ascontiguousarray(base[np.arange(n) % 64])
structs, keep = pkg.make_image_structs(images)
out = np.zeros((n, 1000), np.float32)
rows = (C.POINTER(C.c_float) * n)(*[out[i].ctypes.data_as(C.POINTER(C.c_float)) for i in range(n)])
pin = pkg.PinnedArray(images.shape)
pin.array[...] = images
for rep in range(2):
    t0 = time.perf_counter()
    eng = pkg.Engine(0, 224, prec, max_batch=256)
    t1 = time.perf_counter()
    eng.load_weights(blobs)
    t2 = time.perf_counter()
    pkg._check(pkg.lib().vitb200_forward_structs(eng.h, structs, n, rows))
    t3 = time.perf_counter()
    pkg._check(pkg.lib().vitb200_forward_structs(eng.h, structs, n, rows))
    t3b = time.perf_counter()
    eng.forward_into(pin.array, out)
    t4 = time.perf_counter()
    eng.forward_into(images, out)
    t5 = time.perf_counter()
    eng.close()
    print(f"rep {rep}: create {1e3*(t1-t0):.1f} ms, load_weights {1e3*(t2-t1):.1f} ms, forward_structs({n}) {1e3*(t3-t2):.1f} ms "
          f"= {n/(t3-t2):.0f} img/s (again: {1e3*(t3b-t3):.1f} ms), forward(pinned) {1e3*(t4-t3b):.1f} ms = {n/(t4-t3b):.0f} img/s, "
          f"forward(pageable contiguous) {1e3*(t5-t4):.1f} ms")
