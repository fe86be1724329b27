#!/bin/bash
# W tiles prefetched into the L2 ahead of griddepcontrol.wait (single-CTA GEMM, one work item per CTA): tests + latency A/B
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" 2>&1 | tail -2
python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "fp32 or golden or bf16_engine or bf16_stage" 2>&1 | tail -2
for r in 1 2; do
python tools/b1_latency.py fp32
VITCU_W_PREFETCH=0 python tools/b1_latency.py fp32
python tools/b1_latency.py bf16
VITCU_W_PREFETCH=0 python tools/b1_latency.py bf16
done
