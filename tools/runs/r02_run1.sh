set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_smi.txt
nproc >> gpurun_out/r02a_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?" >> gpurun_out/r02a_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r02a_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_flash -s 13 -c 1 -o gpurun_out/r02a_flash python bench.py --img 384 --batch 64 --steps 2 --warmup 1 --no-extras > gpurun_out/r02a_ncu2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_b1_fp32.csv python tools/b1_forward.py fp32 224 3 > gpurun_out/r02a_ncu3.log 2>&1
timeout 300 python tools/kernel_bench.py 256 20 > gpurun_out/r02a_kernel_bench.json 2>&1
