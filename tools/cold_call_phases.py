#!/usr/bin/env python
"""Phases of repeated COLD ViT_opencl calls (create -> upload -> forward -> destroy) in one process, after the process has
run and released another engine -- the situation of the bench's `dropin.cold` record: python tools/cold_call_phases.py [n] [calls]"""
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
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 5
blobs = pkg.synth.model_blobs(None, 224, seed=0)
with pkg.Engine(0, 224, pkg.BF16, max_batch=256) as e:      # what the bench did before: an engine that ran and was released
    e.load_weights(blobs)
    e.stage(pkg.synth.synthetic_images(256, 224, seed=1))
    e.forward_resident(256)
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
os.environ["VITB200_GPUS"] = "1"
os.environ["VITB200_PERSIST"] = "0"
stats = pkg.CallStats()
alternate = len(sys.argv) > 3 and sys.argv[3] == "alternate"   # staged / plain weight upload in turn, same process
for k in range(calls):
    if alternate:
        os.environ["VITB200_WEIGHT_STAGE"] = str(1 - k % 2)
    t0 = time.perf_counter()
    L.ViT_opencl(imgs, nets, rows)
    dt = time.perf_counter() - t0
    L.vitb200_last_call_stats(C.byref(stats))
    print(f"cold call {k} (weight stage {os.environ.get('VITB200_WEIGHT_STAGE', '1')}): {dt:.3f} s = {n / dt:.0f} images/s (bring-up {stats.create_s:.3f} + weights {stats.weights_s:.3f} + forward {stats.forward_s:.3f} + tear-down {stats.teardown_s:.3f}; library wall {stats.wall_s:.3f})", flush=True)
