#!/bin/bash
# refinement kernel with shared sub-block evaluations + plane kernel tweaks: parity, drop-in, bench
mkdir -p gpurun_out/r2j
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frame.py tests/test_abi.py -m gpu -x -q > gpurun_out/r2j/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j/pytest.log
tail -4 gpurun_out/r2j/pytest.log
timeout 600 python -m pytest tests/test_jm_dropin.py -m gpu -x -q > gpurun_out/r2j/pytest_dropin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j/pytest_dropin.log
tail -4 gpurun_out/r2j/pytest_dropin.log
for c in 2 4; do
  timeout 300 python bench.py --config $c --steps 30 --warmup 3 --no-cpu > gpurun_out/r2j/bench_c$c.json 2> gpurun_out/r2j/bench_c$c.err
done
timeout 300 python bench.py --config 3 --steps 30 --warmup 3 --no-cpu > gpurun_out/r2j/bench_c3.json 2> gpurun_out/r2j/bench_c3.err
python - <<'PY'
import json
for c in (2,3,4):
    try:
        d=json.loads(open(f'gpurun_out/r2j/bench_c{c}.json').read().strip().splitlines()[-1])
        print(c, d['value'], d['e2e']['value'], d.get('kernel_ms_per_step'), d['roofline'].get('worst_case_launch_ms'), d['roofline'].get('worst_case_subpel_refine_ms'))
    except Exception as e: print(c,'ERR',e)
PY
