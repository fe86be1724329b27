#!/bin/bash
# 128 x 64 tiles for the non-summable small-M outputs (BF16 batch-1 qkv / fc1): tests + latency A/B
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" 2>&1 | tail -3
python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "bf16 or golden or variants or 384 or drop_in" 2>&1 | tail -3
for r in 1 2; do
python tools/b1_latency.py bf16
VITCU_GEMM_BN64=0 python tools/b1_latency.py bf16
done
