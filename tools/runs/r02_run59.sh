#!/bin/bash
python -m pytest tests/test_gpu_bench_config_parity.py -x -q -m gpu -k "batch1" -s 2>&1 | tail -12
