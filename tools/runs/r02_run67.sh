#!/bin/bash
# LayerNorm with gamma / beta requested ahead of the reductions (latency-bound launches): A/B, then the LayerNorm + forward tests
tools/ab_run.sh 2 lnlate lnearly -- python tools/b1_latency.py bf16
tools/ab_run.sh 1 lnlate lnearly -- python tools/b1_latency.py fp32
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "layernorm or accumulate" 2>&1 | tail -2
python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "fp32_engine or bf16_engine or stage_by_stage or golden" 2>&1 | tail -2
