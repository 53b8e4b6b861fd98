#!/bin/bash
# tools/gpu_e2e.sh -- end-to-end leg with 1..6 picture streams per GPU
for k in 2 3 4 5 6; do
  python bench.py --steps 60 --warmup 3 --no-cpu --no-worst --e2e-streams $k 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('streams', $k, 'e2e', round(e['value']), 'ms', round(e['ms_per_step'],4), 'value', round(d['value']))"
done
nproc
