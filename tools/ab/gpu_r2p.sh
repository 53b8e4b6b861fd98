#!/bin/bash
# k_epzs_sub at other occupancies (config 3), and the config-4 line again (clock samples)
mkdir -p gpurun_out/r2p
for so in jm_b200/lib/libjmb200.so tools/_bin/libjmb200_ei*.so; do
  JMB200_LIB=$PWD/$so timeout 200 python bench.py --config 3 --steps 20 --warmup 3 --no-cpu --e2e-streams 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$so', 'epzs', round(k['epzs'],4), 'sub', round(k['subpel_refine'],4), 'step', round(d['ms_per_step'],4), 'value', round(d['value']))" | tee -a gpurun_out/r2p/ab.txt
done
timeout 600 python bench.py --config 4 --steps 100 --warmup 3 --cpu-seconds 10 > gpurun_out/r2p/bench_c4.json 2> gpurun_out/r2p/bench_c4.err; echo "bench c4 rc=$?"; cut -c1-200 gpurun_out/r2p/bench_c4.json
