#!/usr/bin/env python
"""A/B of the fc1-shaped launch (M=50432, N=3072, K=768, bf16 out): epilogue BIAS vs BIAS_GELU, with
VITCU_GEMM_EW from the environment -- separates the GELU arithmetic from the rest of the epilogue."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
L = pkg.lib()
pkg.layer_check(L.vitcu_set_device(0))
M = 256 * 197
rng = np.random.default_rng(0)
res = {}
ev0, ev1 = C.c_void_p(), C.c_void_p()
pkg.layer_check(L.vitcu_event_create(C.byref(ev0)))
pkg.layer_check(L.vitcu_event_create(C.byref(ev1)))
for name, N, K, epi in (("fc1_gelu", 3072, 768, pkg.EPI_BIAS_GELU), ("fc1_bias", 3072, 768, pkg.EPI_BIAS),
                        ("qkv_bias", 2304, 768, pkg.EPI_BIAS), ("qkv_gelu", 2304, 768, pkg.EPI_BIAS_GELU),
                        ("sq_bias", 3072, 3072, pkg.EPI_BIAS)):
    a = pkg.DeviceBuffer.from_numpy(pkg.f32_to_bf16_bits(rng.standard_normal((M, K), dtype=np.float32)))
    w = pkg.DeviceBuffer.from_numpy(pkg.f32_to_bf16_bits((rng.standard_normal((N, K), dtype=np.float32) * 0.02).astype(np.float32)))
    bias = pkg.DeviceBuffer.from_numpy(np.zeros(N, np.float32))
    c = pkg.DeviceBuffer(M * N * 2)
    d = pkg.GemmDesc()
    d.M, d.N, d.K, d.lda, d.ldc, d.epilogue, d.out_bf16 = M, N, K, K, N, epi, 1
    d.bias = bias.ptr.value
    for _ in range(3):
        pkg.layer_check(L.vitcu_gemm_bf16(a.ptr, w.ptr, c.ptr, C.byref(d), None))
    pkg.layer_check(L.vitcu_device_sync())
    pkg.layer_check(L.vitcu_event_record(ev0, None))
    for _ in range(20):
        pkg.layer_check(L.vitcu_gemm_bf16(a.ptr, w.ptr, c.ptr, C.byref(d), None))
    pkg.layer_check(L.vitcu_event_record(ev1, None))
    pkg.layer_check(L.vitcu_event_sync(ev1))
    ms = C.c_float()
    pkg.layer_check(L.vitcu_event_elapsed_ms(ev0, ev1, C.byref(ms)))
    res[name] = {"ms": round(ms.value / 20, 4), "tflops": round(2.0 * M * N * K / (ms.value / 20) / 1e9, 1)}
    for b in (a, w, bias, c):
        b.free()
print(json.dumps({"ew": os.environ.get("VITCU_GEMM_EW", "auto"), **res}))
