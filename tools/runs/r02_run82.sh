#!/bin/bash
python bench.py > gpurun_out/bench_s2_final5.json 2> gpurun_out/bench_s2_final5.err; tail -c 200 gpurun_out/bench_s2_final5.err; wc -c gpurun_out/bench_s2_final5.json
