#!/bin/bash
TAG=${1:-r2g}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frame.py -x -q -k "search or frame or surfaces or large" > $O/${TAG}_pytest_search.log 2>&1; echo "pytest search rc=$?"; tail -3 $O/${TAG}_pytest_search.log
bash tools/gpu_variants.sh > $O/${TAG}_variants.txt 2>&1; cat $O/${TAG}_variants.txt
timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
python -c "
import json; d=json.load(open('$O/${TAG}_bench_c3.json')); print(d['value'], d['kernel_ms_per_step'])"
