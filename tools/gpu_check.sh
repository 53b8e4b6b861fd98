#!/bin/bash
# tools/gpu_check.sh -- one gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_smi.txt 2>&1
nproc > $O/${TAG}_nproc.txt; lscpu | head -20 >> $O/${TAG}_nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=12 > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log
tail -5 $O/${TAG}_pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${TAG}_smoke.log
tail -3 $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
cat $O/${TAG}_bench.json; tail -5 $O/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"
cat $O/${TAG}_bench_ref.json; tail -5 $O/${TAG}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > $O/${TAG}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
