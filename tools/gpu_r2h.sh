#!/bin/bash
# tools/gpu_r2h.sh -- EPZS kernels rewritten (cooperative 8x8 Hadamard, parallel selection): parity, then config-3 lines per occupancy variant
TAG=${1:-r2h}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_epzs.py -x -q > $O/${TAG}_pytest_epzs.log 2>&1; echo "pytest epzs rc=$?"; tail -15 $O/${TAG}_pytest_epzs.log
for so in jm_b200/lib/libjmb200.so tools/_bin/libjmb200_ei*.so; do
  JMB200_LIB=$PWD/$so timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$so', 'value', round(d['value']), 'epzs_int ms', round(k['epzs'],3), 'epzs_sub ms', round(k['subpel_refine'],3), 'step', round(d['ms_per_step'],3))"
done
