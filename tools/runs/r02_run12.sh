set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L > gpurun_out/r02l_smi.txt; nproc >> gpurun_out/r02l_smi.txt
timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -x -q -s -k "multi_gpu" > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02l_bench_n2.json 2> gpurun_out/r02l_bench_n2.err; echo "rc=$?" >> gpurun_out/r02l_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r02l_ref_n2.json 2> gpurun_out/r02l_ref_n2.err
