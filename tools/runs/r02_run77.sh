#!/bin/bash
python tools/cold_call_phases.py 4096 10 2>&1 | grep -E "cold call"
python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "reload or persistent or drop_in or fp32_engine or bf16_engine" 2>&1 | tail -2
