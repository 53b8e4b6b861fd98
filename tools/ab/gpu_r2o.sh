#!/bin/bash
# new shared-evaluation parity cases, then the refinement kernel at other occupancies (CTAs/SM via JMB_RF_MINB)
mkdir -p gpurun_out/r2o
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "subpel" 2>&1 | tail -3 | tee gpurun_out/r2o/pytest.txt
for so in jm_b200/lib/libjmb200.so tools/_bin/libjmb200_rf_minb*.so; do
  for mode in "" "--scene-cut"; do
  JMB200_LIB=$PWD/$so timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu --no-worst --e2e-streams 1 $mode 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$so $mode', 'refine', round(k['subpel_refine'],5), 'planes', round(k['subpel_planes'],5), 'step', round(d['ms_per_step'],4), 'value', round(d['value']))" | tee -a gpurun_out/r2o/ab.txt
done; done
