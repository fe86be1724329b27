#!/bin/bash
# three TMA-store staging tiles per epilogue warp vs two (bf16-output GEMMs): correctness, then same-box A/B
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" 2>&1 | tail -2
tools/ab_run.sh 2 bufs2 bufs3 -- python tools/fold_bench.py 256 20
tools/ab_run.sh 2 bufs2 bufs3 -- python bench.py --steps 10 --warmup 3 --no-extras
