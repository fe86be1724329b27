#!/bin/bash
# fused split-bf16 GEMM with one barrier per piece pair: tests + latency
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm or accumulate" 2>&1 | tail -3
python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "fp32 or golden or edge or 384 or variants_match or structs or topk or bf16_engine" 2>&1 | tail -3
for r in 1 2; do
python tools/b1_latency.py fp32
VITCU_FP32_FUSED=0 python tools/b1_latency.py fp32
python tools/b1_latency.py bf16
done
