set -x
cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r02t_plain.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02t_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r02t_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc2 -s 52 -c 4 -o gpurun_out/r02t_gemm4 python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r02t_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_duo -s 13 -c 1 -o gpurun_out/r02t_duo python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r02t_ncu3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02t_b1_fp32.csv python tools/b1_forward.py fp32 224 3 > gpurun_out/r02t_ncu4.log 2>&1
