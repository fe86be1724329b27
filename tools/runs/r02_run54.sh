#!/bin/bash
# split-bf16 GEMM with all pieces of a k-block in one ring slot (six products from tiles loaded once): tests + latency A/B
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "bf16x3 or accumulate" 2>&1 | tail -3
python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "fp32 or golden or edge or 384 or variants_match or structs or topk" 2>&1 | tail -3
for r in 1 2; do
python tools/b1_latency.py fp32
VITCU_FP32_FUSED=0 python tools/b1_latency.py fp32
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/b1_fp32_s2e.csv python tools/b1_forward.py fp32 224 3 > gpurun_out/ncu_b1.log 2>&1
