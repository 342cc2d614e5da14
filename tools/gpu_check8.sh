#!/bin/bash
# small-modulus stage 2 (second iteration): parity, bench, launch list, ncu metrics of the new kernels
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_small.py -x -q -m gpu > gpurun_out/t_small.log 2>&1; echo "small tests rc=$?" >> gpurun_out/summary.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "all gpu tests rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench small rc=$?" >> gpurun_out/summary.txt
MPRES_BENCH_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> gpurun_out/summary.txt
MPRES_BENCH_PROFILER_RANGE=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none -k regex:'k_small_umma|k_align_small|k_ext_norm_small|k_ext_small|k_minplus|k_norm_list|k_outer_info' -c 8 -o gpurun_out/prof_small -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/summary.txt
ncu -i gpurun_out/prof_small.ncu-rep --page raw --csv > gpurun_out/prof_small_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/summary.txt; tail -30 gpurun_out/t_small.log; tail -30 gpurun_out/t_gpu_all.log; cat gpurun_out/bench_small.json; tail -3 gpurun_out/bench_small.err
