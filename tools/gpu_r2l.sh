#!/bin/bash
# tools/gpu_r2l.sh -- whole GPU suite (the packed 8x8 Hadamard now serves k_subpel_refine / k_dist too), config-3 lines per EPZS variant, ncu of k_epzs_sub
TAG=${1:-r2l}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${TAG}_pytest_gpu.log
for so in jm_b200/lib/libjmb200.so tools/_bin/libjmb200_ei*.so; do
  JMB200_LIB=$PWD/$so timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$so', 'value', round(d['value']), 'epzs_int ms', round(k['epzs'],3), 'epzs_sub ms', round(k['subpel_refine'],3), 'step', round(d['ms_per_step'],3))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_epzs_sub -s 1 -c 1 -o $O/prof_${TAG}_epzs_sub python bench.py --config 3 --steps 1 --warmup 2 --no-cpu > $O/prof_${TAG}.log 2>&1; echo "ncu rc=$?"
