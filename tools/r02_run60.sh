#!/bin/bash
# duo attention with one query tile per item at small batch: tests + BF16 batch-1 latency A/B
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention_tensor_core" 2>&1 | tail -3
python -m pytest tests/test_gpu_forward.py tests/test_gpu_bench_config_parity.py -x -q -m gpu -k "bf16_engine or bf16_stage or golden or batch1" 2>&1 | tail -3
for r in 1 2; do
python tools/b1_latency.py bf16
VITCU_ATTN_UNIT_SPLIT=0 python tools/b1_latency.py bf16
done
