#!/bin/bash
# session 2 checkpoint: the whole GPU suite, the default bench line, the reference arm
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py > gpurun_out/bench_s2_1gpu.json 2> gpurun_out/bench_s2_1gpu.err; tail -c 600 gpurun_out/bench_s2_1gpu.err
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/bench_s2_1gpu_20.json 2>/dev/null
