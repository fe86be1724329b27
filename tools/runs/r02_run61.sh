#!/bin/bash
python -m pytest tests/test_gpu_bench_config_parity.py -x -q -m gpu -k "batch1_latency" -s 2>&1 | grep -E "assert|Error|error|batch-1 chain|passed|failed" | head -20
