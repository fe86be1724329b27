#!/bin/bash
# round 2, session 2: attention small v2 (shared K/V buffer, 3 CTAs per SM), accumulate-mode patch embedding at batch 1,
# GEMM exit wait on reads only (A/B against waitw.so = waits for the writes)
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention_simt or accumulate or bf16x3 or gemm_bf16" 2>&1 | tail -5
python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "fp32 or golden or edge or 384 or variants_match or bf16_engine or bf16_stage" 2>&1 | tail -5
tools/ab_run.sh 2 .default waitw -- python tools/b1_latency.py fp32
tools/ab_run.sh 2 .default waitw -- python tools/b1_latency.py bf16
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/b1_fp32_s2b.csv python tools/b1_forward.py fp32 224 3 > gpurun_out/ncu_b1.log 2>&1
