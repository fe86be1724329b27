set -x
cd $GRAFT_REPO_ROOT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc2 -s 8 -c 1 -o gpurun_out/r02o_fc1_fold python tools/fold_bench.py 256 3 fc1 > gpurun_out/r02o_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc2 -s 8 -c 1 -o gpurun_out/r02o_fc2_emit python tools/fold_bench.py 256 3 fc2 > gpurun_out/r02o_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc2 -s 2 -c 1 -o gpurun_out/r02o_fc2_plain python tools/fold_bench.py 256 3 fc2 > gpurun_out/r02o_ncu3.log 2>&1
