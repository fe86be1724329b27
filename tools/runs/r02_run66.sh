#!/bin/bash
# bias requested a chunk ahead in the narrow-tile epilogue: A/B on the batch-1 forwards, then the GEMM tests on the new build
tools/ab_run.sh 2 bias0 bias1 -- python tools/b1_latency.py fp32
tools/ab_run.sh 2 bias0 bias1 -- python tools/b1_latency.py bf16
cp vit-with-opencl_b200/build/ab/bias1.so vit-with-opencl_b200/libvit_b200.so
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm or accumulate" 2>&1 | tail -2
