#!/bin/bash
# deblock / chain kernel changes: parity, deblock line of the bench, 1080p drop-in timing (3 frames)
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_deblock.py tests/test_gpu_frame.py -x -q -k "deblock or chain or mb_surfaces" > $O/ab2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/ab2_pytest.log
timeout 900 python -m pytest tests/test_jm_dropin.py -x -q -k "resident_surfaces or luma_residual_coding_matches" > $O/ab2_pytest_dropin.log 2>&1; echo "pytest dropin rc=$?"; tail -3 $O/ab2_pytest_dropin.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('deblock', d['next_rows']['deblock']); print('value', d['value'], d['kernel_ms_per_step'])"
timeout 900 python tools/dropin_1080p.py ab2 3 > $O/ab2_dropin.log 2>&1; echo "dropin rc=$?"; head -1 $O/ab2_dropin.log | cut -c1-1300
