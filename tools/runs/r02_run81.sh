#!/bin/bash
# ncu --set full of the final batch-1 FP32 kernels: one attention launch, the four GEMM launches of a layer, one LayerNorm
python tools/b1_forward.py fp32 224 3 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"attention_simt_small|gemm_bf16_tc_kernel|layernorm_kernel" -s 120 -c 8 -o gpurun_out/b1_fp32_final -f python tools/b1_forward.py fp32 224 3 > gpurun_out/ncu_b1_final.log 2>&1
tail -2 gpurun_out/ncu_b1_final.log
