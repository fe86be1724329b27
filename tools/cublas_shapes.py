#!/usr/bin/env python
"""Reference point only (not on any product path): what cuBLAS (torch.matmul, bf16) reaches on the four
GEMM shapes of one encoder layer at M = 256*197, plain GEMM without the fused epilogues."""
import torch

M = 256 * 197
for name, N, K in (("qkv", 2304, 768), ("out_proj", 768, 768), ("fc1", 3072, 768), ("fc2", 768, 3072), ("sq", 3072, 3072)):
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        c = a @ w.t()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        c = a @ w.t()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name}: {ms:.4f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s")
