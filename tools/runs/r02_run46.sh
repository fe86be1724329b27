#!/bin/bash
# first-call-in-process cost of the weight upload, staged vs plain (fresh processes, alternating)
for r in 1 2; do
  for st in 1 0; do
    echo "== round $r, VITB200_WEIGHT_STAGE=$st"
    VITB200_WEIGHT_STAGE=$st python tools/persist_timing.py 2>&1 | grep "^ViT_b200: 1024" | head -3
  done
done
