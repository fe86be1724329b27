#!/usr/bin/env python
"""Per-kernel CUDA-event timings at the bench shape (batch*197 rows): the four GEMM launches of a
layer, the attention kernel and LayerNorm.  Usage: python tools/kernel_bench.py [batch] [reps]"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402
import bench  # noqa: E402

pkg = g.load_package()
L = pkg.lib()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T = 197
M = batch * T
pkg.layer_check(L.vitcu_set_device(0))


def timeit(fn):
    ev0, ev1 = C.c_void_p(), C.c_void_p()
    pkg.layer_check(L.vitcu_event_create(C.byref(ev0)))
    pkg.layer_check(L.vitcu_event_create(C.byref(ev1)))
    for _ in range(3):
        fn()
    pkg.layer_check(L.vitcu_device_sync())
    pkg.layer_check(L.vitcu_event_record(ev0, None))
    for _ in range(reps):
        fn()
    pkg.layer_check(L.vitcu_event_record(ev1, None))
    pkg.layer_check(L.vitcu_event_sync(ev1))
    ms = C.c_float()
    pkg.layer_check(L.vitcu_event_elapsed_ms(ev0, ev1, C.byref(ms)))
    return ms.value / reps


out = {"M": M}
per, flops, ms = bench.time_gemms(pkg, L, M, reps)
out["gemm"] = {"layer_ms": ms, "tflops": flops / ms / 1e9, "per_launch": per}

rng = np.random.default_rng(0)
qkv = pkg.DeviceBuffer.from_numpy(pkg.f32_to_bf16_bits(rng.standard_normal((M, 2304), dtype=np.float32)))
att = pkg.DeviceBuffer(M * 768 * 2)
t = timeit(lambda: pkg.layer_check(L.vitcu_attention(qkv.ptr, att.ptr, batch, T, 1, None)))
fl = batch * 12 * 4.0 * T * T * 64
out["attention"] = {"ms": t, "tflops": fl / t / 1e9}

# 577-token attention (384x384 images), 64 images
T2, B2 = 577, 64
qkv2 = pkg.DeviceBuffer.from_numpy(pkg.f32_to_bf16_bits(rng.standard_normal((B2 * T2, 2304), dtype=np.float32)))
att2 = pkg.DeviceBuffer(B2 * T2 * 768 * 2)
t = timeit(lambda: pkg.layer_check(L.vitcu_attention(qkv2.ptr, att2.ptr, B2, T2, 1, None)))
out["attention_577x64"] = {"ms": t, "tflops": B2 * 12 * 4.0 * T2 * T2 * 64 / t / 1e9}

x = pkg.DeviceBuffer.from_numpy(rng.standard_normal((M, 768), dtype=np.float32))
gam = pkg.DeviceBuffer.from_numpy(np.ones(768, np.float32))
bet = pkg.DeviceBuffer.from_numpy(np.zeros(768, np.float32))
y = pkg.DeviceBuffer(M * 768 * 2)
t = timeit(lambda: pkg.layer_check(L.vitcu_layernorm(x.ptr, 768, y.ptr, 1, gam.ptr, bet.ptr, M, None)))
out["layernorm_bf16"] = {"ms": t, "gbs": M * 768 * 6 / t / 1e6}
assert L.vitcu_watchdog_check() == 0
print(json.dumps({k: v for k, v in out.items() if k != "gemm"}))
print(json.dumps(out["gemm"]))

# the MEASURED_PEAKS.json shape (cuBLAS 8192^3): mainloop-dominated, epilogue negligible
if os.environ.get("KB_SQUARE", "1") == "1":
    n = 8192
    a = pkg.DeviceBuffer.from_numpy(pkg.f32_to_bf16_bits(rng.standard_normal((n, n), dtype=np.float32)))
    w = pkg.DeviceBuffer.from_numpy(pkg.f32_to_bf16_bits(rng.standard_normal((n, n), dtype=np.float32) * np.float32(0.01)))
    bias = pkg.DeviceBuffer.from_numpy(np.zeros(n, np.float32))
    c = pkg.DeviceBuffer(n * n * 2)
    d = pkg.GemmDesc()
    d.M, d.N, d.K, d.lda, d.ldc, d.epilogue, d.out_bf16 = n, n, n, n, n, pkg.EPI_BIAS, 1
    d.bias = bias.ptr.value
    t = timeit(lambda: pkg.layer_check(L.vitcu_gemm_bf16(a.ptr, w.ptr, c.ptr, C.byref(d), None)))
    print(json.dumps({"square_8192": {"ms": t, "tflops": 2.0 * n * n * n / t / 1e9}}))
