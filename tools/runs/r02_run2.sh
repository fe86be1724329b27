set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention_tensor_core" > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
for r in 1 2; do
  VITCU_ATTN_KERNEL=solo timeout 120 python tools/attn_ab.py >> gpurun_out/r02b_ab.log 2>&1
  VITCU_ATTN_KERNEL=duo timeout 120 python tools/attn_ab.py >> gpurun_out/r02b_ab.log 2>&1
done
timeout 600 python -m pytest tests/test_gpu_bench_config_parity.py -m gpu -x -q -s -k bf16_batch256 >> gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r02b_bench.json 2>gpurun_out/r02b_bench.err
VITCU_ATTN_KERNEL=solo timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r02b_bench_solo.json 2>>gpurun_out/r02b_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_duo -s 13 -c 1 -o gpurun_out/r02b_duo python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r02b_ncu.log 2>&1
