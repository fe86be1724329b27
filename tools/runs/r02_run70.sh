#!/bin/bash
# patch embedding with bias / position rows requested a chunk ahead: kernel alone at 256 images, batch-1 forwards, tests
tools/ab_run.sh 2 pe_now pe_ahead -- python tools/pe_bench.py 256 20
tools/ab_run.sh 2 pe_now pe_ahead -- python tools/b1_latency.py bf16
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "patch_embed" 2>&1 | tail -2
python -m pytest tests/test_gpu_forward.py tests/test_gpu_bench_config_parity.py -x -q -m gpu -k "bf16" 2>&1 | tail -2
