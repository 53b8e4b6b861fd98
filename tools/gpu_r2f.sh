#!/bin/bash
# tools/gpu_r2f.sh -- two-level gate of k_int_search: parity suite, then A/B lines of the launch-shape variants; config 4 bench
TAG=${1:-r2f}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frame.py -x -q -k "search or frame or surfaces or large or subpel" > $O/${TAG}_pytest_search.log 2>&1; echo "pytest search rc=$?"; tail -6 $O/${TAG}_pytest_search.log
bash tools/gpu_variants.sh > $O/${TAG}_variants.txt 2>&1; cat $O/${TAG}_variants.txt
timeout 600 python bench.py --config 4 --steps 30 --warmup 3 --cpu-seconds 8 > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err; echo "bench c4 rc=$?"
cat $O/${TAG}_bench_c4.json | cut -c1-4000; tail -5 $O/${TAG}_bench_c4.err
