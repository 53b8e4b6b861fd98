#!/bin/bash
# tools/gpu_r2m.sh -- jmb_mb_chain: parity test, drop-in tests (bitstream identity with the run-ahead on and off), 1080p drop-in timing; config-3 line
TAG=${1:-r2m}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_frame.py -x -q -k "chain or mb_surfaces" > $O/${TAG}_pytest_chain.log 2>&1; echo "pytest chain rc=$?"; tail -15 $O/${TAG}_pytest_chain.log
timeout 900 python -m pytest tests/test_jm_dropin.py -x -q -k "resident_surfaces or bitstream" > $O/${TAG}_pytest_dropin.log 2>&1; echo "pytest dropin rc=$?"; tail -15 $O/${TAG}_pytest_dropin.log
timeout 900 python tools/dropin_1080p.py $TAG 3 > $O/${TAG}_dropin_1080p.log 2>&1; echo "dropin rc=$?"; tail -4 $O/${TAG}_dropin_1080p.log | cut -c1-1800
JMB_SHIM_CHAIN=0 timeout 900 python tools/dropin_1080p.py ${TAG}_nochain 3 > $O/${TAG}_dropin_1080p_nochain.log 2>&1; echo "dropin nochain rc=$?"; tail -4 $O/${TAG}_dropin_1080p_nochain.log | cut -c1-1200
timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('c3 value', round(d['value']), 'epzs_int ms', round(k['epzs'],3), 'epzs_sub ms', round(k['subpel_refine'],3), 'step', round(d['ms_per_step'],3))"
