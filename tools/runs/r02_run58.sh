#!/bin/bash
# end of session 2: whole GPU suite, smoke, default bench line, launch list of the bench command
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py > gpurun_out/bench_s2_final.json 2> gpurun_out/bench_s2_final.err; tail -c 300 gpurun_out/bench_s2_final.err
python bench.py --steps 2 --warmup 1 --no-extras > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s2_final_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
