set -x
cd $GRAFT_REPO_ROOT
python tools/fold_bench.py 256 20 > gpurun_out/r02k_fold.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "layernorm_fold or gemm_bf16" > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log
timeout 900 python -m pytest tests/test_gpu_bench_config_parity.py tests/test_gpu_forward.py -m gpu -x -q -s > gpurun_out/r02k_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest2.log
for r in 1 2; do
  VITB200_LN_FOLD=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras >> gpurun_out/r02k_bench_nofold.json 2>>gpurun_out/r02k_bench.err
  timeout 300 python bench.py --steps 20 --warmup 5 --no-extras >> gpurun_out/r02k_bench_fold.json 2>>gpurun_out/r02k_bench.err
done
