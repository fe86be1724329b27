#!/usr/bin/env python
"""ONE persistent ViT_opencl call over 4096 pageable images on all GPUs of the box, for several staging-slot sizes and
copy-thread counts: python tools/dropin_slots.py [images]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
L = pkg.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
gpus = pkg.device_count()
blobs = pkg.synth.model_blobs(None, 224, seed=0)
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
os.environ["VITB200_GPUS"] = str(gpus)
for rnd in range(2):
    for mb, th in ((4, 8), (16, 8), (16, 4), (32, 8)):
        os.environ["VITB200_STAGE_SLOT_MB"] = str(mb)
        os.environ["VITB200_STAGE_THREADS"] = str(th)
        best = 1e9
        for it in range(5):
            t0 = time.perf_counter()
            L.ViT_opencl(imgs, nets, rows)
            dt = time.perf_counter() - t0
            if it:
                best = min(best, dt)
        L.vitb200_release_persistent()
        print(f"round {rnd}: {gpus} GPUs, slot {mb} MB, {th} copy threads per GPU: {1e3 * best:.1f} ms = {n / best:.0f} images/s", flush=True)
