set -x
cd $GRAFT_REPO_ROOT
bash tools/ab_run.sh 1 emit_v0 emit_v1 emit_v2 emit_v3 -- python tools/fold_bench.py 256 20 > gpurun_out/r02h_fold.log 2>&1
