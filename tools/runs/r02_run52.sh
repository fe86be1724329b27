#!/bin/bash
# 8 GPUs: strong scaling of ONE ViT_opencl call (config 4), streaming stores on / off
python tools/dropin_scaling.py 4096 2>&1 | tail -1
VITB200_STAGE_STREAMING=0 python tools/dropin_scaling.py 4096 2>&1 | tail -1
