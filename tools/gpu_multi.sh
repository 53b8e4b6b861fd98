#!/bin/bash
# tools/gpu_multi.sh TAG N -- the N-GPU bench lines (run under `gpurun --gpus N`): config 2 (independent picture streams) and
# config 5 (every rank codes a picture against rank 0's reconstructed anchor, read over NVLink)
TAG=${1:-r02}; N=${2:-2}
O=gpurun_out; mkdir -p $O
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > $O/${TAG}_$2_n$N.json 2> $O/${TAG}_$2_n$N.err; echo "$2 n$N rc=$?"; tail -1 $O/${TAG}_$2_n$N.json | cut -c1-700; tail -3 $O/${TAG}_$2_n$N.err; }
run 29521 bench_c2 --steps 50 --warmup 3
run 29522 bench_c5 --config 5 --steps 20 --warmup 3
run 29523 bench_c3 --config 3 --steps 20 --warmup 3 --no-cpu
