set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention_tensor_core" > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
bash tools/ab_run.sh 2 duo_p00 duo_p11 duo_p55 -- python tools/attn_ab.py > gpurun_out/r02e_ab.log 2>&1
cp vit-with-opencl_b200/build/ab/duo_p55.so vit-with-opencl_b200/libvit_b200.so
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention_tensor_core" > gpurun_out/r02e_pytest_p55.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest_p55.log
