cd $GRAFT_REPO_ROOT
bash tools/ab_run.sh 2 ctl_park ctl_spin -- python tools/attn_ab.py 577 64 > gpurun_out/r02ab_flash.log 2>&1
bash tools/ab_run.sh 2 ctl_park ctl_spin -- python tools/attn_ab.py 197 256 > gpurun_out/r02ab_duo.log 2>&1
