#!/bin/bash
# tools/gpu_r2c.sh -- surfaces/arg-min + EPZS (split kernels) tests, drop-in bitstream tests, config-3 bench, 1080p drop-in timing
TAG=${1:-r2c}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_epzs.py tests/test_gpu_frame.py -x -q > $O/${TAG}_pytest_new.log 2>&1; echo "pytest new rc=$?"; tail -12 $O/${TAG}_pytest_new.log
timeout 400 python bench.py --config 3 --steps 20 --warmup 3 --no-cpu > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
cat $O/${TAG}_bench_c3.json | cut -c1-1800; tail -5 $O/${TAG}_bench_c3.err
timeout 1500 python -m pytest tests/test_jm_dropin.py -m gpu -q --durations=8 > $O/${TAG}_pytest_dropin.log 2>&1; echo "pytest dropin rc=$?"; tail -25 $O/${TAG}_pytest_dropin.log
cat $O/dropin_times.txt
timeout 900 python tools/dropin_1080p.py $TAG 3 > $O/${TAG}_dropin_1080p.log 2>&1; echo "dropin 1080p rc=$?"; tail -5 $O/${TAG}_dropin_1080p.log | cut -c1-2500
