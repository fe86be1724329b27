cd $GRAFT_REPO_ROOT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_flash_duo -s 3 -c 1 -o gpurun_out/r02aa_fd3 python tools/attn_ab.py 577 64 > gpurun_out/r02aa_ncu.log 2>&1
