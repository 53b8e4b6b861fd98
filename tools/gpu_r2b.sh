#!/bin/bash
# tools/gpu_r2b.sh -- EPZS parity + frame-form tests, bench lines of configs 2 and 3
TAG=${1:-r2b}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_epzs.py tests/test_gpu_frame.py -x -q > $O/${TAG}_pytest_new.log 2>&1; echo "pytest new rc=$?"; tail -15 $O/${TAG}_pytest_new.log
timeout 400 python bench.py --config 3 --steps 20 --warmup 3 > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
cat $O/${TAG}_bench_c3.json | cut -c1-3500; tail -5 $O/${TAG}_bench_c3.err
timeout 400 python bench.py --steps 30 --warmup 3 --no-cpu > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; echo "bench c2 rc=$?"
cat $O/${TAG}_bench_c2.json | cut -c1-2500; tail -5 $O/${TAG}_bench_c2.err
