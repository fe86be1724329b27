cd $GRAFT_REPO_ROOT
bash tools/ab_run.sh 2 emit_tma emit_lsu -- python tools/fold_bench.py 256 20 > gpurun_out/r02ad_fold.log 2>&1
cp vit-with-opencl_b200/build/ab/emit_lsu.so vit-with-opencl_b200/libvit_b200.so
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fold or e4m3" > gpurun_out/r02ad_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ad_pytest.log
