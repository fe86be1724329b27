set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l > gpurun_out/r02ae_smi.txt; nproc >> gpurun_out/r02ae_smi.txt
timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -x -q -k "multi_gpu" > gpurun_out/r02ae_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ae_pytest.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02ae_bench_n8.json 2> gpurun_out/r02ae_bench_n8.err; echo "rc=$?" >> gpurun_out/r02ae_bench_n8.err
python tools/dropin_scaling.py 4096 2>&1 | grep -v "^ViT_b200" > gpurun_out/r02ae_dropin.log
