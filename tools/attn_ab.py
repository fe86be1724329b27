#!/usr/bin/env python
"""Event-timed attention launches (B=256, T=197), `reps` launches per sample, several samples; run it
under different VITCU_ATTN_* settings in one gpurun call to A/B kernel variants on the same box."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
L = pkg.lib()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 197
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rng = np.random.default_rng(0)
qkv = pkg.DeviceBuffer.from_numpy(pkg.f32_to_bf16_bits(rng.standard_normal((B * T, 2304), dtype=np.float32)))
out = pkg.DeviceBuffer(B * T * 768 * 2)
ev0, ev1 = C.c_void_p(), C.c_void_p()
pkg.layer_check(L.vitcu_event_create(C.byref(ev0)))
pkg.layer_check(L.vitcu_event_create(C.byref(ev1)))
res = []
for sample in range(5):
    for _ in range(5):
        pkg.layer_check(L.vitcu_attention(qkv.ptr, out.ptr, B, T, 1, None))
    pkg.layer_check(L.vitcu_device_sync())
    pkg.layer_check(L.vitcu_event_record(ev0, None))
    for _ in range(50):
        pkg.layer_check(L.vitcu_attention(qkv.ptr, out.ptr, B, T, 1, None))
    pkg.layer_check(L.vitcu_event_record(ev1, None))
    pkg.layer_check(L.vitcu_event_sync(ev1))
    ms = C.c_float()
    pkg.layer_check(L.vitcu_event_elapsed_ms(ev0, ev1, C.byref(ms)))
    res.append(ms.value / 50 * 1e3)
print({k: v for k, v in os.environ.items() if k.startswith("VITCU_ATTN")}, "us per launch:", " ".join(f"{r:.1f}" for r in res))
