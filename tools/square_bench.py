#!/usr/bin/env python
"""One mainloop-dominated GEMM (n^3, bias epilogue) for profiling the tcgen05 pipeline."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
L = pkg.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rng = np.random.default_rng(0)
a = pkg.DeviceBuffer.from_numpy(pkg.f32_to_bf16_bits(rng.standard_normal((n, n), dtype=np.float32)))
w = pkg.DeviceBuffer.from_numpy(pkg.f32_to_bf16_bits(rng.standard_normal((n, n), dtype=np.float32) * np.float32(0.01)))
bias = pkg.DeviceBuffer.from_numpy(np.zeros(n, np.float32))
c = pkg.DeviceBuffer(n * n * 2)
d = pkg.GemmDesc()
d.M, d.N, d.K, d.lda, d.ldc, d.epilogue, d.out_bf16 = n, n, n, n, n, pkg.EPI_BIAS, 1
d.bias = bias.ptr.value
for _ in range(reps):
    pkg.layer_check(L.vitcu_gemm_bf16(a.ptr, w.ptr, c.ptr, C.byref(d), None))
pkg.layer_check(L.vitcu_device_sync())
print("done")
