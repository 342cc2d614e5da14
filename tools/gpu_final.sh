#!/bin/bash
# round-end record: smoke, full GPU suite, default bench line (e2e + cpu_baseline), reference arm, config 2/4/5 lines, launch list, ncu full summary
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "all gpu tests rc=$?" >> gpurun_out/summary.txt
( time timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?" >> gpurun_out/summary.txt
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?" >> gpurun_out/summary.txt
for w in gemm1024_106bit gemm2048_106bit gemm2048_212bit gemm2048_318bit gemm2048_424bit gemm2048_530bit gemm2048_636bit gemm2048_742bit gemm2048_848bit; do
  timeout 600 python bench.py --workload $w --no-e2e --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?" >> gpurun_out/summary.txt
done
timeout 600 python bench.py --full-precision-inputs --workload gemm1024_106bit --no-e2e --no-cpu-baseline > gpurun_out/bench_gemm1024_fullprec.json 2> gpurun_out/bench_gemm1024_fullprec.err; echo "bench fullprec rc=$?" >> gpurun_out/summary.txt
for w in gemv16384_212bit gemvt16384_212bit dot16m_212bit; do
  timeout 600 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?" >> gpurun_out/summary.txt
done
MPRES_BENCH_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> gpurun_out/summary.txt
MPRES_BENCH_PROFILER_RANGE=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none -k regex:'k_small_umma|k_align_small|k_ext_small|k_norm_fast|k_mp_gather|k_outer_info' -c 12 -o gpurun_out/prof_final -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/summary.txt
ncu -i gpurun_out/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/summary.txt; cat gpurun_out/smoke.log | tail -2; tail -4 gpurun_out/t_gpu_all.log | cut -c1-200; cut -c1-300 gpurun_out/bench_default.json; tail -4 gpurun_out/bench_default.err
