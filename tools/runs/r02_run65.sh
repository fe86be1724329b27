#!/bin/bash
# 8 GPUs under torchrun, final build of the session: the complete bench line incl. config5 / fp8 / dropin
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_s2_8gpu.json 2> gpurun_out/bench_s2_8gpu.err
tail -c 300 gpurun_out/bench_s2_8gpu.err; wc -c gpurun_out/bench_s2_8gpu.json
