#!/bin/bash
# tools/gpu_ab.sh -- parity of the search kernels, then A/B bench lines of the k_int_search builds in tools/_bin against the library (2 repetitions, normal regime + scene-cut probe)
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frame.py -x -q -k "search or frame or surfaces or large or scattered" > $O/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/ab_pytest.log
for rep in 1 2; do
for so in jm_b200/lib/libjmb200.so tools/_bin/libjmb200_nt*.so; do
  for mode in "" "--scene-cut"; do
  JMB200_LIB=$PWD/$so python bench.py --steps 30 --warmup 3 --no-cpu --e2e-streams 1 $mode 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$so $mode', 'int_search ms', round(d['kernel_ms_per_step']['int_search'],4), 'step', round(d['ms_per_step'],4), 'value', round(d['value']))"
done; done; done
