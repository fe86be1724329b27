#!/bin/bash
# ncu --set full of one batch-1 FP32 attention launch and one accumulate-mode qkv GEMM launch
ncu --set full --clock-control none --import-source on -k regex:attention_simt_small -s 14 -c 1 -o gpurun_out/b1_attn_small -f python tools/b1_forward.py fp32 224 3 > gpurun_out/ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc_kernel -s 60 -c 4 -o gpurun_out/b1_gemm -f python tools/b1_forward.py fp32 224 3 > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_attn.log gpurun_out/ncu_gemm.log
