#!/bin/bash
# tools/gpu_r2j.sh -- EPZS kernels: parity, config-3 lines per variant, then ncu --set full of both EPZS kernels (default build)
TAG=${1:-r2j}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_epzs.py -x -q > $O/${TAG}_pytest_epzs.log 2>&1; echo "pytest epzs rc=$?"; tail -5 $O/${TAG}_pytest_epzs.log
for so in jm_b200/lib/libjmb200.so tools/_bin/libjmb200_ei*.so; do
  JMB200_LIB=$PWD/$so timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$so', 'value', round(d['value']), 'epzs_int ms', round(k['epzs'],3), 'epzs_sub ms', round(k['subpel_refine'],3), 'step', round(d['ms_per_step'],3))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_epzs -s 2 -c 2 -o $O/prof_${TAG}_epzs python bench.py --config 3 --steps 1 --warmup 2 --no-cpu > $O/prof_${TAG}.log 2>&1; echo "ncu rc=$?"
