cd $GRAFT_REPO_ROOT
export VITCU_PDL=0
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention_tensor_core and duo and (197-3 or 129-1 or 50-2) or attention_flash_tensor_core and duo and (577-2 or 257-1) or fold_producer and 6304 or fold_consumer and 6500 or e4m3 and (6304-3072 or 6400-768) or flash_score_ranges and duo and 12" > gpurun_out/r02af_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02af_memcheck.log
