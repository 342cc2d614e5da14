#!/bin/bash
# parity of the new stage-3 kernels, bench, launch list and ncu captures of the three top kernels
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_blas.py -x -q -m gpu -k "stage or reduced" > gpurun_out/t_stage3.log 2>&1; echo "stage3 tests rc=$?" >> gpurun_out/summary.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "all gpu tests rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?" >> gpurun_out/summary.txt
MPRES_BENCH_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> gpurun_out/summary.txt
MPRES_BENCH_PROFILER_RANGE=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_limb_umma|k_norm_fast|k_align_planes|k_minplus' -c 5 -o gpurun_out/prof_r2 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -15 gpurun_out/t_stage3.log; tail -3 gpurun_out/t_gpu_all.log; cat gpurun_out/bench_r2.json
