#!/bin/bash
python bench.py > gpurun_out/bench_s2_final4.json 2> gpurun_out/bench_s2_final4.err; tail -c 200 gpurun_out/bench_s2_final4.err
python -m pytest tests/test_gpu_bench_config_parity.py tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -2
