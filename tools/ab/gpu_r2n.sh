#!/bin/bash
# k_subpel_planes launch shapes (warps per CTA x rows per thread): plane parity and the kernel's time at 1080p / 4K for each build
mkdir -p gpurun_out/r2n
for so in tools/_bin/libjmb200_sp_nw*.so; do
  JMB200_LIB=$PWD/$so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "subpel_planes or frame_search_with_subpel" 2>&1 | tail -1 | tee -a gpurun_out/r2n/ab.txt
  for mode in "" "--config 3"; do
  JMB200_LIB=$PWD/$so timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu --no-worst --e2e-streams 1 $mode 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$so $mode', 'planes', round(k['subpel_planes'],5), 'step', round(d['ms_per_step'],4), 'value', round(d['value']))" | tee -a gpurun_out/r2n/ab.txt
done; done
