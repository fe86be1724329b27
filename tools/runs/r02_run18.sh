set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; echo "rc=$?" >> gpurun_out/r02r_bench.err
