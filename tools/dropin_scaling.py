#!/usr/bin/env python
"""Strong scaling of ONE ViT_opencl call (BASELINE config 4: 4096 images, BF16, persistent context) over 1..N GPUs of the
box, single process -- the product's own split (vit_opencl.c).  python tools/dropin_scaling.py [images]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
L = pkg.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
blobs = pkg.synth.model_blobs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "Network"), 224, seed=0)
base = pkg.synth.synthetic_images(64, 224, seed=4096)
bufs = [np.array(base[i % 64], dtype=np.float32, order="C", copy=True) for i in range(n)]
imgs = (pkg.ImageData * n)()
for i, b in enumerate(bufs):
    imgs[i].n, imgs[i].c, imgs[i].h, imgs[i].w = n, 3, 224, 224
    imgs[i].data = b.ctypes.data_as(C.POINTER(C.c_float))
nets, keep = pkg.make_network_structs(blobs)
out = np.zeros((n, 1000), np.float32)
rows = (C.POINTER(C.c_float) * n)(*[out[i].ctypes.data_as(C.POINTER(C.c_float)) for i in range(n)])
os.environ["VITB200_PRECISION"] = "bf16"
os.environ["VITB200_PERSIST"] = "1"
res = {}
ref = None
for gpus in [k for k in (1, 2, 4, 8) if k <= pkg.device_count()]:
    os.environ["VITB200_GPUS"] = str(gpus)
    best = 1e9
    for it in range(5):
        t0 = time.perf_counter()
        L.ViT_opencl(imgs, nets, rows)
        dt = time.perf_counter() - t0
        if it:
            best = min(best, dt)
    if ref is None:
        ref = out.copy()
    res[f"gpus{gpus}"] = {"s": round(best, 4), "images_per_s": round(n / best), "rows_equal": bool(np.array_equal(out, ref))}
L.vitb200_release_persistent()
print(json.dumps(res))
