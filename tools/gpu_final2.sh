#!/bin/bash
# tools/gpu_final2.sh TAG -- end-of-round refresh after the sub-pel kernel work: whole GPU suite, smoke, bench lines of configs 2, 3, 4
# (the reference arms and the drop-in encoder timing of tools/gpu_final.sh are unchanged by kernel work and are not repeated),
# ncu launch list + (FULL_CAPTURE=1) --set full captures of the kernels that changed, sanitizer passes over their tests
TAG=${1:-r02}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest_gpu.log; tail -4 $O/${TAG}_pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench c2 rc=$?"; cut -c1-600 $O/${TAG}_bench.json
timeout 600 python bench.py --config 3 --steps 20 --warmup 3 > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; echo "bench c3 rc=$?"; cut -c1-300 $O/${TAG}_bench_c3.json
timeout 600 python bench.py --config 4 --steps 100 --warmup 3 --cpu-seconds 10 > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err; echo "bench c4 rc=$?"; cut -c1-300 $O/${TAG}_bench_c4.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/${TAG}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
[ -n "$FULL_CAPTURE" ] && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mc_tq|k_subpel_planes|k_subpel_refine|k_gen_requests|k_pack_results' -s 10 -c 5 -o $O/prof_${TAG}_subpel python bench.py --steps 1 --warmup 2 --no-cpu --no-worst > $O/prof_${TAG}_c.log 2>&1; echo "ncu sub-pel kernels rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_epzs.py tests/test_gpu_frame.py -x -q -k "subpel or frame_search or epzs_frame or u8_uploads" > $O/${TAG}_sanitizer_memcheck2.log 2>&1; echo "memcheck rc=$?" | tee -a $O/${TAG}_sanitizer_memcheck2.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frame.py -x -q -k "shared_evaluations or frame_search_with_subpel or subpel_planes or u8_uploads" > $O/${TAG}_sanitizer_racecheck2.log 2>&1; echo "racecheck rc=$?" | tee -a $O/${TAG}_sanitizer_racecheck2.log
echo done
