#!/bin/bash
# k_mc_tq_modes_c after the quantiser-pass change (4x4 two-pass, 8x8 one-pass): parity, config-3 time, one ncu capture of the 4x4 kernel
mkdir -p gpurun_out/r2s
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frame.py -m gpu -x -q > gpurun_out/r2s/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s/pytest.log
tail -3 gpurun_out/r2s/pytest.log
timeout 300 python bench.py --config 3 --steps 40 --warmup 3 --no-cpu > gpurun_out/r2s/bench_c3.json 2> gpurun_out/r2s/bench_c3.err
python -c "
import json
d=json.loads(open('gpurun_out/r2s/bench_c3.json').read().strip().splitlines()[-1])
print(3, round(d['value']), round(d['e2e']['value']), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items() if v})"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mc_tq -s 2 -c 1 -o gpurun_out/r2s/mc_tq python bench.py --steps 1 --warmup 2 --no-cpu --no-worst > gpurun_out/r2s/ncu.log 2>&1; echo "ncu rc=$?"
