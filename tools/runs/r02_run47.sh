#!/bin/bash
# host-bound drop-in call on one GPU (1 and 2 staging threads): tail split and streaming stores, A/B
for th in 1 2; do
for r in 1 2; do
python tools/dropin_hostbound.py 512 $th 2>&1 | tail -1
VITB200_TAIL_SPLIT=0 python tools/dropin_hostbound.py 512 $th 2>&1 | tail -1
VITB200_STAGE_STREAMING=0 python tools/dropin_hostbound.py 512 $th 2>&1 | tail -1
done
done
python tools/dropin_hostbound.py 4096 8 2>&1 | tail -1
VITB200_TAIL_SPLIT=0 VITB200_STAGE_STREAMING=0 python tools/dropin_hostbound.py 4096 8 2>&1 | tail -1
