#!/bin/bash
# k_mc_tq_modes_c with the cheaper quantiser passes: parity (4x4 and 8x8, standard and table scans), bench lines of configs 2 / 3 / 4
mkdir -p gpurun_out/r2r
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frame.py tests/test_abi.py -m gpu -x -q > gpurun_out/r2r/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r/pytest.log
tail -3 gpurun_out/r2r/pytest.log
for c in 2 3 4; do
  timeout 300 python bench.py --config $c --steps 40 --warmup 3 --cpu-seconds 4 --no-worst > gpurun_out/r2r/bench_c$c.json 2> gpurun_out/r2r/bench_c$c.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2r/bench_c$c.json').read().strip().splitlines()[-1])
print($c, round(d['value']), round(d['e2e']['value']), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items() if v}, d['cpu_baseline'].get('checked') or d['cpu_baseline'].get('gpu_matches_reference_on_sample'))
PY
done
