#!/bin/bash
# requests formed inside the kernels (working tree) against the request array + generator / packer launches (HEAD), same box, 2 repetitions
mkdir -p gpurun_out/r2t
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frame.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/r2t/pytest.txt
for rep in 1 2; do for so in tools/_bin/libjmb200_head.so jm_b200/lib/libjmb200.so; do
  JMB200_LIB=$PWD/$so timeout 200 python bench.py --steps 60 --warmup 3 --no-cpu --no-worst --e2e-streams 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$so', {a:round(b,4) for a,b in k.items() if b}, 'step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))" | tee -a gpurun_out/r2t/ab.txt
done; done
