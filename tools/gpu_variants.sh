#!/bin/bash
# bench line of every tuning build in tools/_bin (normal and, with arg "cut", the scene-cut probe)
for so in tools/_bin/libjmb200_nt*.so; do
  for mode in "" "--scene-cut"; do
  JMB200_LIB=$PWD/$so python bench.py --steps 10 --warmup 3 --no-cpu $mode 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$so $mode', 'int_search ms', round(d['kernel_ms_per_step']['int_search'],4), 'step', round(d['ms_per_step'],4))"
  done
done
