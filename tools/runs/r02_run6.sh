set -x
cd $GRAFT_REPO_ROOT
bash tools/ab_run.sh 2 duo_p00 duo_p01 duo_p11 duo_p49 duo_p11r2 -- python tools/attn_ab.py > gpurun_out/r02f_ab.log 2>&1
