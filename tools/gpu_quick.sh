#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_small.py -x -q -m gpu > gpurun_out/t_small.log 2>&1; echo "small tests rc=$?" >> gpurun_out/summary.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "all gpu tests rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench small rc=$?" >> gpurun_out/summary.txt
MPRES_BENCH_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -30 gpurun_out/t_small.log | cut -c1-200; tail -8 gpurun_out/t_gpu_all.log | cut -c1-200; cut -c1-300 gpurun_out/bench_small.json; tail -3 gpurun_out/bench_small.err
