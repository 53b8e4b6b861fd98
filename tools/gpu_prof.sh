#!/bin/bash
# tools/gpu_prof.sh TAG [kernel-regex] -- parity tests, bench line, then one ncu --set full capture of the kernel
TAG=${1:-p}; K=${2:-k_int_search}
bash tools/gpu_quick.sh $TAG
ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log | cut -c1-200
