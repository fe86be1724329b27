cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "flash" > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest.log
for r in 1 2; do
VITCU_ATTN_KERNEL=solo timeout 120 python tools/attn_ab.py 577 64 >> gpurun_out/r02z_ab.log 2>&1
VITCU_ATTN_KERNEL=duo timeout 120 python tools/attn_ab.py 577 64 >> gpurun_out/r02z_ab.log 2>&1
done
