cd $GRAFT_REPO_ROOT
bash tools/ab_run.sh 2 fd_p00 fd_p01 fd_p11 -- python tools/attn_ab.py 577 64 > gpurun_out/r02w_ab.log 2>&1
