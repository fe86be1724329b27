#!/bin/bash
# staged weight upload: forward tests, then cold-call breakdown with and without staging
python -m pytest tests/test_gpu_forward.py -x -q -m gpu 2>&1 | tail -3
python tools/dropin_breakdown.py 2048 bf16 2>&1 | grep rep
VITB200_WEIGHT_STAGE=0 python tools/dropin_breakdown.py 2048 bf16 2>&1 | grep rep
python tools/persist_timing.py 2>&1 | grep -v "^ViT_b200: weight"
