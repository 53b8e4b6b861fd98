#!/bin/bash
# bench line of every tuning build in tools/_bin (3 repetitions each, normal regime; last one also the scene-cut probe)
for rep in 1 2 3; do
for so in tools/_bin/libjmb200_nt*.so; do
  JMB200_LIB=$PWD/$so python bench.py --steps 30 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$so', 'int_search ms', round(d['kernel_ms_per_step']['int_search'],4), 'step', round(d['ms_per_step'],4))"
done; done
