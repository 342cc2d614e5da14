#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_small.py -x -q -m gpu > gpurun_out/t_small.log 2>&1; echo "small tests rc=$?" >> gpurun_out/summary.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "all gpu tests rc=$?" >> gpurun_out/summary.txt
for v in small small_k64; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --stage2 $v > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?" >> gpurun_out/summary.txt
done
for w in gemv16384_212bit gemvt16384_212bit dot16m_212bit gemm1024_106bit; do
timeout 600 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?" >> gpurun_out/summary.txt
done
timeout 300 python bench.py --impl reference --workload dot16m_212bit --steps 2 --warmup 1 > gpurun_out/bench_ref_dot.json 2> gpurun_out/bench_ref_dot.err
MPRES_BENCH_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> gpurun_out/summary.txt
MPRES_BENCH_PROFILER_RANGE=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none -k regex:'k_small_umma|k_align_small|k_ext_small|k_norm_fast' -c 5 -o gpurun_out/prof_small -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/summary.txt
ncu -i gpurun_out/prof_small.ncu-rep --page raw --csv > gpurun_out/prof_small_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/summary.txt; tail -30 gpurun_out/t_small.log | cut -c1-200; tail -8 gpurun_out/t_gpu_all.log | cut -c1-200; for f in gpurun_out/bench_*.json; do echo $f; cut -c1-250 $f; done; tail -3 gpurun_out/bench_gemv16384_212bit.err
