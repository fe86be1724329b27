#!/bin/bash
# attention small v3 (18-query tiles, wavefront-efficient layouts): tests, latency, launch list
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention_simt" 2>&1 | tail -3
python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "fp32 or golden or edge or variants_match" 2>&1 | tail -3
python tools/b1_latency.py fp32
VITCU_ATTN_SMALL=0 python tools/b1_latency.py fp32
python tools/b1_latency.py fp32
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/b1_fp32_s2c.csv python tools/b1_forward.py fp32 224 3 > gpurun_out/ncu_b1.log 2>&1
grep attention_simt_small gpurun_out/b1_fp32_s2c.csv | tail -3
