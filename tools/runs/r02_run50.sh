#!/bin/bash
# head split of the first chunk: end-to-end A/B (bench e2e record), then the forward tests that compare chunkings
for r in 1 2; do
for hs in 1 0; do
echo "== round $r head_split=$hs"
VITB200_HEAD_SPLIT=$hs python bench.py --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'single', round(d['e2e']['single_batch_call_images_per_s']))"
done
done
python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "pageable or 4096 or full_batch or structs or drop_in or persistent" 2>&1 | tail -3
