#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "all gpu tests rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_nb6.json 2> gpurun_out/bench_nb6.err; echo "bench nb6 rc=$?" >> gpurun_out/summary.txt
for v in; do
MPRES_B200_LIB=$PWD/build_ab/libmpres_b200_$v.so timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?" >> gpurun_out/summary.txt
done
MPRES_BENCH_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -8 gpurun_out/t_gpu_all.log | cut -c1-200; for f in gpurun_out/bench_nb*.json; do echo $f; cut -c1-200 $f; done; tail -3 gpurun_out/bench_nb6.err
