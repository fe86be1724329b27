#!/bin/bash
# probe: what does launching the single-CTA GEMM as clusters of 2 / 4 independent CTAs cost?
for r in 1 2; do
python tools/b1_latency.py bf16
VITCU_PROBE_CLUSTER=2 python tools/b1_latency.py bf16
VITCU_PROBE_CLUSTER=4 python tools/b1_latency.py bf16
done
