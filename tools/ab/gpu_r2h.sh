#!/bin/bash
# staging change of the request generators / result packers: parity + the three bench configs
mkdir -p gpurun_out/r2h
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h/pytest.log
for c in 2 3 4; do
  timeout 300 python bench.py --config $c --steps 30 --warmup 3 --no-cpu > gpurun_out/r2h/bench_c$c.json 2> gpurun_out/r2h/bench_c$c.err
done
tail -3 gpurun_out/r2h/pytest.log
python - <<'PY'
import json
for c in (2,3,4):
    try:
        d=json.loads(open(f'gpurun_out/r2h/bench_c{c}.json').read().strip().splitlines()[-1])
        print(c, d['value'], d['e2e']['value'], d.get('kernel_ms_per_step'))
    except Exception as e: print(c,'ERR',e)
PY
