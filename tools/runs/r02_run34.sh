cd $GRAFT_REPO_ROOT
bash tools/ab_run.sh 2 duo_narrow duo_wide -- python tools/attn_ab.py 197 256 > gpurun_out/r02ah_ab.log 2>&1
cp vit-with-opencl_b200/build/ab/duo_wide.so vit-with-opencl_b200/libvit_b200.so
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention_tensor_core and duo" > gpurun_out/r02ah_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ah_pytest.log
