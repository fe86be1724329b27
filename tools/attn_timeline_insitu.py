#!/usr/bin/env python
"""Period of the attention kernel's units (clock64 stamps) inside a real forward (eager, VITB200_NO_GRAPH=1)
vs launched alone: separates clock effects from cycle effects."""
import ctypes as C
import os
import sys

import numpy as np

os.environ["VITB200_NO_GRAPH"] = "1"
os.environ.setdefault("VITCU_ATTN_DBG_FIRST", "20")
os.environ.setdefault("VITCU_ATTN_DBG_CTA", "70")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
L = pkg.lib()
L.vitcu_attention_debug_timeline.argtypes = [C.c_void_p]
B, T = 256, 197
dbg = pkg.DeviceBuffer(4 * 16 * 8 * 8)


def periods():
    t = dbg.to_numpy(np.uint64, (4, 16, 8)).astype(np.int64)
    pub = t[0, :, 5]  # left exp warp group: P published
    pub = pub[pub > 0]
    return np.diff(pub)


blobs = pkg.synth.model_blobs(None, 224, seed=7)
x = pkg.synth.synthetic_images(B, 224, seed=1)
with pkg.Engine(0, 224, pkg.BF16, max_batch=B) as e:
    e.load_weights(blobs)
    e.stage(x)
    for _ in range(10):
        e.forward_resident(B)
    pkg.layer_check(L.vitcu_memset(dbg.ptr, 0, 4 * 16 * 8 * 8, None))
    L.vitcu_attention_debug_timeline(dbg.ptr)
    ms = e.forward_resident(B)
    L.vitcu_attention_debug_timeline(None)
    d = periods()
    print(f"in situ (last layer of an eager forward, {ms:.2f} ms): unit period cycles median {np.median(d):.0f}  mean {d.mean():.0f}  {d.tolist()}")
rng = np.random.default_rng(0)
qkv = pkg.DeviceBuffer.from_numpy(pkg.f32_to_bf16_bits(rng.standard_normal((B * T, 2304), dtype=np.float32)))
out = pkg.DeviceBuffer(B * T * 768 * 2)
for _ in range(3):
    pkg.layer_check(L.vitcu_attention(qkv.ptr, out.ptr, B, T, 1, None))
pkg.layer_check(L.vitcu_memset(dbg.ptr, 0, 4 * 16 * 8 * 8, None))
L.vitcu_attention_debug_timeline(dbg.ptr)
pkg.layer_check(L.vitcu_attention(qkv.ptr, out.ptr, B, T, 1, None))
L.vitcu_attention_debug_timeline(None)
d = periods()
print(f"alone: unit period cycles median {np.median(d):.0f}  mean {d.mean():.0f}  {d.tolist()}")
