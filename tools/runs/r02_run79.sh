#!/bin/bash
# head gemv with the weight row requested in one go: A/B on the batch-1 forwards + its tests
tools/ab_run.sh 2 base gemv -- python tools/b1_latency.py bf16
tools/ab_run.sh 1 base gemv -- python tools/b1_latency.py fp32
cp vit-with-opencl_b200/build/ab/gemv.so vit-with-opencl_b200/libvit_b200.so
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py -x -q -m gpu -k "sgemm or fp32_engine or bf16_engine or golden or topk" 2>&1 | tail -2
