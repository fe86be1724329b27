cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "flash" > gpurun_out/r02ac_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ac_pytest.log
VITCU_FD_SHAPE=64 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "flash and duo" >> gpurun_out/r02ac_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ac_pytest.log
for r in 1 2; do
VITCU_ATTN_KERNEL=solo timeout 120 python tools/attn_ab.py 577 64 >> gpurun_out/r02ac_ab.log 2>&1
timeout 120 python tools/attn_ab.py 577 64 >> gpurun_out/r02ac_ab.log 2>&1
VITCU_FD_SHAPE=64 timeout 120 python tools/attn_ab.py 577 64 >> gpurun_out/r02ac_ab.log 2>&1
done
timeout 900 python -m pytest tests -m gpu -x -q -k "384 or variants" >> gpurun_out/r02ac_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ac_pytest.log
