#!/bin/bash
# duo attention with one query tile per item at small batch + K-sliced TF32 patch embedding: tests + BF16 batch-1 latency A/B
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention_tensor_core or patch_embed" 2>&1 | tail -3
python -m pytest tests/test_gpu_forward.py tests/test_gpu_bench_config_parity.py -x -q -m gpu -k "bf16 or golden or batch1 or variants" 2>&1 | tail -3
for r in 1 2; do
python tools/b1_latency.py bf16
VITCU_ATTN_UNIT_SPLIT=0 python tools/b1_latency.py bf16
VITB200_PE_SPLITK=0 python tools/b1_latency.py bf16
VITCU_ATTN_UNIT_SPLIT=0 VITB200_PE_SPLITK=0 python tools/b1_latency.py bf16
done
