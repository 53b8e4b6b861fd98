#!/bin/bash
# tools/gpu_r2n.sh -- deblocking kernel: parity vs the oracle, live-encoder verification, bitstream identity with DeblockFrame on the device; chain test; smoke
TAG=${1:-r2n}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_deblock.py tests/test_gpu_frame.py -x -q -k "deblock or chain" > $O/${TAG}_pytest_deblock.log 2>&1; echo "pytest deblock rc=$?"; tail -12 $O/${TAG}_pytest_deblock.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/${TAG}_smoke.log
timeout 1500 python -m pytest tests/test_jm_dropin.py -x -q > $O/${TAG}_pytest_dropin.log 2>&1; echo "pytest dropin rc=$?"; tail -12 $O/${TAG}_pytest_dropin.log
