#!/bin/bash
for k in ${@:-1 2 3 4}; do python bench.py --steps 60 --warmup 3 --no-cpu --e2e-streams $k 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams $k', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['e2e']['ms_per_step'],3), d['e2e']['host_ms_per_picture'])"; done
