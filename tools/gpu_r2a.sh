#!/bin/bash
# tools/gpu_r2a.sh -- round-2 first check: GPU tests (new + old), bench line of config 2, scene-cut probe
TAG=${1:-r2a}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_frame.py -x -q > $O/${TAG}_pytest_frame.log 2>&1; echo "pytest frame rc=$?"; tail -15 $O/${TAG}_pytest_frame.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_frame.py > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/${TAG}_pytest_gpu.log
timeout 400 python bench.py --steps 30 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
cat $O/${TAG}_bench.json; tail -5 $O/${TAG}_bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --scene-cut > $O/${TAG}_bench_scenecut.json 2> $O/${TAG}_bench_scenecut.err; echo "scenecut rc=$?"
cat $O/${TAG}_bench_scenecut.json | head -c 1500
