#!/bin/bash
# end-to-end leg with 1..8 picture streams per GPU (config 2 and 3)
mkdir -p gpurun_out/r2q; nproc | tee gpurun_out/r2q/ab.txt
for c in 2 3; do for n in 2 3 4 6 8; do
  timeout 200 python bench.py --config $c --steps 20 --warmup 3 --no-cpu --no-worst --e2e-streams $n 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('config $c streams $n', 'e2e', round(e['value']), 'value', round(d['value']), {k:v for k,v in e.items() if 'stream' in k or 'single' in k})" | tee -a gpurun_out/r2q/ab.txt
done; done
