set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ai_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02ai_smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02ai_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ai_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02ai_bench.json 2> gpurun_out/r02ai_bench.err; echo "rc=$?" >> gpurun_out/r02ai_bench.err
timeout 900 python bench.py > gpurun_out/r02ai_bench_default.json 2>> gpurun_out/r02ai_bench.err; echo "rc=$?" >> gpurun_out/r02ai_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02ai_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r02ai_ncu1.log 2>&1
