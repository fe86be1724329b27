#!/bin/bash
# BF16 batch-1 launch list
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/b1_bf16_s2.csv python tools/b1_forward.py bf16 224 3 > gpurun_out/ncu_b1.log 2>&1
tail -2 gpurun_out/ncu_b1.log
