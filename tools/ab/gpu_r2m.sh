#!/bin/bash
# shared sub-block evaluations in k_subpel_refine (third pass) and k_epzs_sub: parity, drop-in EPZS, A/B against the previous commit
mkdir -p gpurun_out/r2m
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frame.py tests/test_gpu_epzs.py tests/test_epzs_golden.py tests/test_abi.py -m gpu -x -q > gpurun_out/r2m/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m/pytest.log
tail -4 gpurun_out/r2m/pytest.log
timeout 400 python -m pytest tests/test_jm_dropin.py -m gpu -x -q -k "bundled" > gpurun_out/r2m/pytest_dropin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m/pytest_dropin.log
tail -3 gpurun_out/r2m/pytest_dropin.log
for so in tools/_bin/libjmb200_old_refine.so jm_b200/lib/libjmb200.so; do
  for mode in "" "--scene-cut" "--config 3"; do
  JMB200_LIB=$PWD/$so timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu --no-worst --e2e-streams 1 $mode 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$so $mode', 'refine ms', round(k['subpel_refine'],4), 'int', round(k['int_search'],4), 'epzs', round(k['epzs'],4), 'step', round(d['ms_per_step'],4), 'value', round(d['value']))" | tee -a gpurun_out/r2m/ab.txt
done; done
