#!/bin/bash
# tools/gpu_quick.sh -- GPU parity tests + one bench line (no CPU legs); every step under its own timeout
TAG=${1:-q}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${TAG}_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
cat $O/${TAG}_bench.json; tail -5 $O/${TAG}_bench.err
