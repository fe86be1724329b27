#!/bin/bash
# end of session 2 (after the BF16 batch-1 changes): whole GPU suite, smoke, default bench line
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py > gpurun_out/bench_s2_final2.json 2> gpurun_out/bench_s2_final2.err; tail -c 300 gpurun_out/bench_s2_final2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/b1_bf16_s2b.csv python tools/b1_forward.py bf16 224 3 > gpurun_out/ncu_b1.log 2>&1
