#!/bin/bash
# tools/gpu_r2d.sh -- chroma parity + live-encoder verification, fused per-partition search, 1080p drop-in timing
TAG=${1:-r2d}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_chroma.py tests/test_gpu_frame.py -x -q > $O/${TAG}_pytest_new.log 2>&1; echo "pytest new rc=$?"; tail -12 $O/${TAG}_pytest_new.log
timeout 900 python -m pytest tests/test_jm_dropin.py -m gpu -q -k "verify or live_encoder or resident_surfaces or epzs_subpelgrid or full_search_baseline or fast_full" > $O/${TAG}_pytest_dropin.log 2>&1; echo "pytest dropin rc=$?"; tail -25 $O/${TAG}_pytest_dropin.log
timeout 600 python tools/dropin_1080p.py $TAG 3 > $O/${TAG}_dropin_1080p.log 2>&1; echo "dropin 1080p rc=$?"; tail -5 $O/${TAG}_dropin_1080p.log | cut -c1-1800
