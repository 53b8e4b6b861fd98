#!/bin/bash
# k_subpel_planes rewrite: plane parity (both upload paths), the picture-form tests, bench configs 2 and 3, one ncu capture
mkdir -p gpurun_out/r2i
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frame.py -m gpu -x -q > gpurun_out/r2i/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i/pytest.log
tail -4 gpurun_out/r2i/pytest.log
for c in 2 3; do
  timeout 300 python bench.py --config $c --steps 30 --warmup 3 --no-cpu > gpurun_out/r2i/bench_c$c.json 2> gpurun_out/r2i/bench_c$c.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_subpel_planes -s 3 -c 1 -o gpurun_out/r2i/subpel_planes python bench.py --steps 2 --warmup 1 --no-cpu --no-worst > gpurun_out/r2i/ncu.log 2>&1
python - <<'PY'
import json
for c in (2,3):
    try:
        d=json.loads(open(f'gpurun_out/r2i/bench_c{c}.json').read().strip().splitlines()[-1])
        print(c, d['value'], d['e2e']['value'], d.get('kernel_ms_per_step'))
    except Exception as e: print(c,'ERR',e)
PY
