#!/bin/bash
# zero-fill after the row loads, head gemv / softmax with their loads up front, tensor maps fetched before the PDL wait: A/B + tests
tools/ab_run.sh 2 lnearly tails -- python tools/b1_latency.py bf16
tools/ab_run.sh 2 lnearly tails -- python tools/b1_latency.py fp32
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "layernorm or sgemm or softmax or accumulate or attention_tensor_core" 2>&1 | tail -2
python -m pytest tests/test_gpu_forward.py tests/test_gpu_bench_config_parity.py -x -q -m gpu -k "fp32_engine or bf16_engine or stage_by_stage or golden or batch1_latency or topk or edge" 2>&1 | tail -2
