#!/bin/bash
# tools/gpu_final.sh TAG -- the measurements of a round on ONE B200: whole GPU suite, smoke, bench lines of configs 2 (both arms), 3, 4,
# ncu launch list + --set full captures of the dominant kernels, sanitizer passes, drop-in encoder timing, pipe-rate microbenchmark
TAG=${1:-r02}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log; tail -4 $O/${TAG}_pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench c2 rc=$?"; cut -c1-600 $O/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "ref rc=$?"; cut -c1-400 $O/${TAG}_bench_reference.json
timeout 600 python bench.py --config 3 --steps 20 --warmup 3 > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; echo "bench c3 rc=$?"; cut -c1-300 $O/${TAG}_bench_c3.json
timeout 600 python bench.py --config 3 --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_c3_reference.json 2> $O/${TAG}_bench_c3_reference.err; echo "ref c3 rc=$?"
timeout 600 python bench.py --config 4 --steps 30 --warmup 3 --cpu-seconds 10 > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err; echo "bench c4 rc=$?"; cut -c1-300 $O/${TAG}_bench_c4.json
timeout 600 python bench.py --config 4 --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_c4_reference.json 2> $O/${TAG}_bench_c4_reference.err; echo "ref c4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/${TAG}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_int_search -s 2 -c 1 -o $O/prof_${TAG}_int_search python bench.py --steps 1 --warmup 3 --no-cpu --no-worst > $O/prof_${TAG}_a.log 2>&1; echo "ncu int_search rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_epzs_int|k_epzs_sub' -s 2 -c 2 -o $O/prof_${TAG}_epzs python bench.py --config 3 --steps 1 --warmup 2 --no-cpu > $O/prof_${TAG}_b.log 2>&1; echo "ncu epzs rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_deblock|k_mc_tq|k_subpel_planes|k_subpel_refine' -s 8 -c 4 -o $O/prof_${TAG}_other python bench.py --steps 1 --warmup 2 --no-cpu --no-worst > $O/prof_${TAG}_c.log 2>&1; echo "ncu other rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_epzs.py tests/test_gpu_deblock.py tests/test_gpu_frame.py -x -q -k "full_search_random or frame_search or epzs_search_matches or deblock_matches or chain or fast_full_search" > $O/${TAG}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/${TAG}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_epzs.py tests/test_gpu_deblock.py tests/test_gpu_frame.py -x -q -k "epzs_frame or deblock_matches or chain or frame_search_with_subpel" > $O/${TAG}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/${TAG}_sanitizer_racecheck.log
timeout 1200 python tools/dropin_1080p.py ${TAG} 5 > $O/${TAG}_dropin_1080p.log 2>&1; echo "dropin rc=$?"; tail -3 $O/${TAG}_dropin_1080p.log | cut -c1-900
[ -x tools/_bin/ubench ] && timeout 120 tools/_bin/ubench > $O/${TAG}_ubench.txt 2>&1
echo done
