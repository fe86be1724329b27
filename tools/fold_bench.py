#!/usr/bin/env python
"""CUDA-event timings of the GEMM launches of one encoder layer with and without the folded LayerNorm
(M = batch * 197): qkv / fc1 as consumers (ln_stats epilogue), out-proj / fc2 as producers (emit epilogue),
and the layernorm kernel they replace.  python tools/fold_bench.py [batch] [reps]"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
L = pkg.lib()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
only = sys.argv[3] if len(sys.argv) > 3 else None
M, D = batch * 197, 768
pkg.layer_check(L.vitcu_set_device(0))
rng = np.random.default_rng(0)
ev0, ev1 = C.c_void_p(), C.c_void_p()
pkg.layer_check(L.vitcu_event_create(C.byref(ev0)))
pkg.layer_check(L.vitcu_event_create(C.byref(ev1)))


def timeit(fn):
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
    return ms.value / reps * 1e3


def bf(shape, scale=1.0):
    return pkg.DeviceBuffer.from_numpy(pkg.f32_to_bf16_bits((rng.standard_normal(shape, dtype=np.float32) * scale).astype(np.float32)))


slots = D // 128
x = pkg.DeviceBuffer.from_numpy(rng.standard_normal((M, D), dtype=np.float32))
xb = pkg.DeviceBuffer(M * D * 2)
st = pkg.DeviceBuffer(slots * M * 8)
pkg.layer_check(L.vitcu_rowstats_cast(x.ptr, xb.ptr, st.ptr, M, D, slots, None))
out = {"M": M}
for name, N, K, epi, kind in (("qkv", 2304, 768, pkg.EPI_BIAS, "consumer"), ("out_proj", 768, 768, pkg.EPI_BIAS_RESIDUAL, "producer"),
                              ("fc1", 3072, 768, pkg.EPI_BIAS_GELU, "consumer"), ("fc2", 768, 3072, pkg.EPI_BIAS_RESIDUAL, "producer")):
    if only and name != only:
        continue
    a = xb if K == 768 and kind == "consumer" else bf((M, K))
    w = bf((N, K), 0.02)
    bias = pkg.DeviceBuffer.from_numpy(np.zeros(N, np.float32))
    cs = pkg.DeviceBuffer.from_numpy(np.zeros(N, np.float32))
    res = {}
    for fold in (0, 1):
        d = pkg.GemmDesc()
        d.M, d.N, d.K, d.lda, d.ldc, d.epilogue = M, N, K, K, N, epi
        d.bias = bias.ptr.value
        if kind == "consumer":
            c = pkg.DeviceBuffer(M * N * 2)
            d.out_bf16 = 1
            if fold:
                d.ln_stats, d.ln_slots, d.ln_colsum = st.ptr.value, slots, cs.ptr.value
        else:
            c = x
            d.residual = x.ptr.value
            if fold:
                d.emit_bf16, d.emit_stats = xb.ptr.value, st.ptr.value
        us = timeit(lambda: pkg.layer_check(L.vitcu_gemm_bf16(a.ptr, w.ptr, c.ptr, C.byref(d), None)))
        res["fold" if fold else "plain"] = {"us": round(us, 1), "tflops": round(2.0 * M * N * K / us / 1e6, 1)}
    out[name] = res
g_ = pkg.DeviceBuffer.from_numpy(np.ones(D, np.float32))
out["layernorm_us"] = round(timeit(lambda: pkg.layer_check(L.vitcu_layernorm(x.ptr, D, xb.ptr, 1, g_.ptr, g_.ptr, M, None))), 1)
out["rowstats_cast_us"] = round(timeit(lambda: pkg.layer_check(L.vitcu_rowstats_cast(x.ptr, xb.ptr, st.ptr, M, D, slots, None))), 1)
print(json.dumps(out))
