cd $GRAFT_REPO_ROOT
bash tools/ab_run.sh 2 flash_old flash_p00 flash_p11 -- python tools/flash_ab.py > gpurun_out/r02u_flash_ab.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "flash" > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02u_pytest.log
