#!/bin/bash
# A/B of the refinement kernel: previous commit's library vs the working tree, normal content and the unrelated-reference probe
mkdir -p gpurun_out/r2k
for so in tools/_bin/libjmb200_old_refine.so jm_b200/lib/libjmb200.so; do
  for mode in "" "--scene-cut"; do
  JMB200_LIB=$PWD/$so timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu --no-worst --e2e-streams 1 $mode 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$so $mode', 'refine ms', round(k['subpel_refine'],4), 'int', round(k['int_search'],4), 'planes', round(k['subpel_planes'],4), 'step', round(d['ms_per_step'],4), 'value', round(d['value']))" | tee -a gpurun_out/r2k/ab.txt
done; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_subpel_refine -s 3 -c 1 -o gpurun_out/r2k/subpel_refine python bench.py --steps 2 --warmup 1 --no-cpu --no-worst > gpurun_out/r2k/ncu.log 2>&1
