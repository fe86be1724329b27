cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02aj_bench.json 2> gpurun_out/r02aj_bench.err; echo "rc=$?" >> gpurun_out/r02aj_bench.err
