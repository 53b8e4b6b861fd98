#!/bin/bash
# EPZS picture form with the requests formed inside the kernels: EPZS / picture-form parity, the config-3 line, one capture of the final config-2 kernels
mkdir -p gpurun_out/r2u
timeout 200 python -m pytest tests/test_gpu_epzs.py tests/test_epzs_golden.py tests/test_gpu_frame.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/r2u/pytest.txt
timeout 200 python bench.py --config 3 --steps 20 --warmup 3 > gpurun_out/r2u/bench_c3.json 2> gpurun_out/r2u/bench_c3.err; echo "c3 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2u/bench_c3.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['e2e']['value']), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items() if v}, d['cpu_baseline'].get('checked') or d['cpu_baseline'].get('gpu_matches_reference_on_sample'))"
true
