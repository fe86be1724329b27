set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "layernorm_fold or gemm_bf16" > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log
timeout 900 python -m pytest tests/test_gpu_bench_config_parity.py tests/test_gpu_forward.py -m gpu -x -q -s > gpurun_out/r02g_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest2.log
for r in 1 2; do
  VITB200_LN_FOLD=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras >> gpurun_out/r02g_bench_nofold.json 2>>gpurun_out/r02g_bench.err
  timeout 300 python bench.py --steps 20 --warmup 5 --no-extras >> gpurun_out/r02g_bench_fold.json 2>>gpurun_out/r02g_bench.err
done
