cd $GRAFT_REPO_ROOT
python tools/dropin_scaling.py 4096 2>&1 | grep -v "^ViT_b200" > gpurun_out/r02s_dropin.log
