set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L > gpurun_out/r02q_smi.txt; nproc >> gpurun_out/r02q_smi.txt; free -g >> gpurun_out/r02q_smi.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02q_bench_n4.json 2> gpurun_out/r02q_bench_n4.err; echo "rc=$?" >> gpurun_out/r02q_bench_n4.err
