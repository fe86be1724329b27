#!/bin/bash
# 2 GPUs under torchrun, final build of the session: the complete bench line incl. config5 / fp8 / dropin
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_s2_2gpu.json 2> gpurun_out/bench_s2_2gpu.err
tail -c 400 gpurun_out/bench_s2_2gpu.err; wc -c gpurun_out/bench_s2_2gpu.json
