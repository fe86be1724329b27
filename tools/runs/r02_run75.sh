#!/bin/bash
# one device arena for the activation buffers: cold-call phases, then the forward tests
VITB200_DEBUG_TEARDOWN=1 python tools/cold_call_phases.py 4096 10 2>&1 | grep -E "cold call|tear-down:" | grep -v "^ViT_b200 tear-down: sync 0.0000, device frees 0.0[0-4]"
python -m pytest tests/test_gpu_forward.py -x -q -m gpu 2>&1 | tail -2
