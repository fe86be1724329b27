#!/bin/bash
# round 2, session 2: FP32 batch-1 chain (accumulate-mode qkv / fc1, cp.async attention) -- tests, then same-box A/B
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention_simt or accumulate or split3 or layernorm or bf16x3" 2>&1 | tail -5
python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "fp32 or golden or edge or 384 or variants_match" 2>&1 | tail -5
for r in 1 2; do
  python tools/b1_latency.py fp32
  VITB200_FP32_SPLITK=0 python tools/b1_latency.py fp32
  VITCU_ATTN_SMALL=0 python tools/b1_latency.py fp32
  VITB200_FP32_SPLITK=0 VITCU_ATTN_SMALL=0 python tools/b1_latency.py fp32
done
python tools/b1_latency.py bf16
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/b1_fp32_s2.csv python tools/b1_forward.py fp32 224 3 > gpurun_out/ncu_b1.log 2>&1
