#!/usr/bin/env python
"""Print the per-unit timeline (SM clock cycles) of CTA 0 of the single-block attention kernel."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
L = pkg.lib()
B, T = 256, 197
rng = np.random.default_rng(0)
qkv = pkg.DeviceBuffer.from_numpy(pkg.f32_to_bf16_bits(rng.standard_normal((B * T, 2304), dtype=np.float32)))
out = pkg.DeviceBuffer(B * T * 768 * 2)
dbg = pkg.DeviceBuffer(4 * 16 * 8 * 8)
for _ in range(2):
    pkg.layer_check(L.vitcu_attention(qkv.ptr, out.ptr, B, T, 1, None))
pkg.layer_check(L.vitcu_memset(dbg.ptr, 0, 4 * 16 * 8 * 8, None))
L.vitcu_attention_debug_timeline.argtypes = [C.c_void_p]
L.vitcu_attention_debug_timeline(dbg.ptr)
pkg.layer_check(L.vitcu_attention(qkv.ptr, out.ptr, B, T, 1, None))
L.vitcu_attention_debug_timeline(None)
t = dbg.to_numpy(np.uint64, (4, 16, 8)).astype(np.int64)
t0 = t[t > 0].min()
names = ["wait S", "S ready", "-", "max seen", "exp done", "P pub", "-"]
for role, rn in ((0, "left WG"), (1, "right WG")):
    print(rn, "(cycles since start; columns:", ", ".join(names), ")")
    for k in range(12):
        row = t[role, k, :7]
        print(f"  unit {k:2d}: " + " ".join(f"{(v - t0) if v else -1:7d}" for v in row) +
              "   | deltas " + " ".join(f"{(row[i + 1] - row[i]) if row[i] and row[i + 1] else -1:6d}" for i in range(6)))
print("MMA issuer (S issued, P seen, PV issued)")
for k in range(12):
    row = t[2, k, :3]
    print(f"  unit {k:2d}: " + " ".join(f"{(v - t0) if v else -1:7d}" for v in row))
print("statistics/epilogue warps (wait S, row max published, epilogue of the same unit issued)")
for k in range(12):
    row = t[3, k, [0, 1, 6]]
    print(f"  unit {k:2d}: " + " ".join(f"{(v - t0) if v else -1:7d}" for v in row))
