cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -x -q -k "large_chunks" > gpurun_out/r02ag_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ag_pytest.log
