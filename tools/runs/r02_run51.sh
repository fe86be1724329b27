#!/bin/bash
# 2 GPUs: the product's own split (test) after restricting chunk splits to pair-eligible quarter chunks
python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "multi_gpu or 4096 or pageable" 2>&1 | tail -2
