#!/bin/bash
# tools/gpu_r2e.sh (2 GPUs) -- chroma tests, bench lines of config 4 (N=1) and config 5 (N=2, anchor over NVLink), config 2 at N=2
TAG=${1:-r2e}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_chroma.py -x -q > $O/${TAG}_pytest_chroma.log 2>&1; echo "pytest chroma rc=$?"; tail -6 $O/${TAG}_pytest_chroma.log
timeout 600 python bench.py --config 4 --steps 30 --warmup 3 --cpu-seconds 8 > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err; echo "bench c4 rc=$?"
cat $O/${TAG}_bench_c4.json | cut -c1-3800; tail -5 $O/${TAG}_bench_c4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config 5 --steps 20 --warmup 3 > $O/${TAG}_bench_c5_n2.json 2> $O/${TAG}_bench_c5_n2.err; echo "bench c5 n2 rc=$?"
cat $O/${TAG}_bench_c5_n2.json | cut -c1-2500; tail -5 $O/${TAG}_bench_c5_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 3 > $O/${TAG}_bench_c2_n2.json 2> $O/${TAG}_bench_c2_n2.err; echo "bench c2 n2 rc=$?"
cat $O/${TAG}_bench_c2_n2.json | cut -c1-1800; tail -5 $O/${TAG}_bench_c2_n2.err
