#!/bin/bash
# L2 prefetch of weight-like vectors ahead of griddepcontrol.wait (LayerNorm gamma / beta, head rows, position rows, conv filter boxes): A/B + tests
tools/ab_run.sh 2 lnearly pf -- python tools/b1_latency.py bf16
tools/ab_run.sh 2 lnearly pf -- python tools/b1_latency.py fp32
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "layernorm or patch_embed or sgemm or softmax" 2>&1 | tail -2
python -m pytest tests/test_gpu_forward.py tests/test_gpu_bench_config_parity.py -x -q -m gpu -k "fp32_engine or bf16_engine or stage_by_stage or golden or batch1_latency" 2>&1 | tail -2
