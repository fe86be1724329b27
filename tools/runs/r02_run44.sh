#!/bin/bash
# split-K on CTA pairs for the small-M reduce-add GEMMs: tests, then latency A/B against the single-CTA kernel
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" 2>&1 | tail -3
python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "fp32 or golden or edge or variants_match or bf16_engine or bf16_stage or 384" 2>&1 | tail -3
for r in 1 2; do
python tools/b1_latency.py fp32
VITCU_GEMM_MODE=1cta python tools/b1_latency.py fp32
python tools/b1_latency.py bf16
VITCU_GEMM_MODE=1cta python tools/b1_latency.py bf16
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/b1_fp32_s2d.csv python tools/b1_forward.py fp32 224 3 > gpurun_out/ncu_b1.log 2>&1
