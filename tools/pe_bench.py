#!/usr/bin/env python
"""CUDA-event timing of the TF32 patch embedding alone (the kernel's launches back to back): python tools/pe_bench.py [batch] [reps]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
L = pkg.lib()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
pkg.layer_check(L.vitcu_set_device(0))
rng = np.random.default_rng(0)
img = pkg.DeviceBuffer.from_numpy(rng.standard_normal((batch, 3, 224, 224), dtype=np.float32))
w = pkg.DeviceBuffer.from_numpy((rng.standard_normal((768, 768), dtype=np.float32) * 0.03).astype(np.float32))
b = pkg.DeviceBuffer.from_numpy(rng.standard_normal(768, dtype=np.float32))
pos = pkg.DeviceBuffer.from_numpy(rng.standard_normal((197, 768), dtype=np.float32))
x = pkg.DeviceBuffer(batch * 197 * 768 * 4)
ev0, ev1 = C.c_void_p(), C.c_void_p()
pkg.layer_check(L.vitcu_event_create(C.byref(ev0)))
pkg.layer_check(L.vitcu_event_create(C.byref(ev1)))
for _ in range(3):
    pkg.layer_check(L.vitcu_patch_embed_tc(img.ptr, w.ptr, b.ptr, pos.ptr, x.ptr, batch, 224, None))
pkg.layer_check(L.vitcu_device_sync())
pkg.layer_check(L.vitcu_event_record(ev0, None))
for _ in range(reps):
    pkg.layer_check(L.vitcu_patch_embed_tc(img.ptr, w.ptr, b.ptr, pos.ptr, x.ptr, batch, 224, None))
pkg.layer_check(L.vitcu_event_record(ev1, None))
pkg.layer_check(L.vitcu_event_sync(ev1))
ms = C.c_float()
pkg.layer_check(L.vitcu_event_elapsed_ms(ev0, ev1, C.byref(ms)))
print(f"patch_embed_tc, {batch} images: {1e3 * ms.value / reps:.1f} us per launch")
