#!/bin/bash
# full GPU check after a re-entry: parity tests, default bench line (e2e + cpu_baseline), reference arm,
# config-4 vector probe, launch list, ncu full captures of the GEMM stages and the vector kernels
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "all gpu tests rc=$?" >> gpurun_out/summary.txt
( time timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?" >> gpurun_out/summary.txt
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?" >> gpurun_out/summary.txt
VEC_CFGS=0 timeout 600 python tools/vec_probe.py 16 16384 16384 > gpurun_out/vec_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/summary.txt
MPRES_BENCH_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> gpurun_out/summary.txt
MPRES_BENCH_PROFILER_RANGE=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_limb_umma|k_norm_fast|k_align_planes4|k_minplus|k_norm_list|k_base_extend|k_outer_info' -c 9 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/summary.txt
VEC_CFGS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_mv_acc' -c 12 -o gpurun_out/prof_vec -f python tools/vec_probe.py 16 8192 8192 4194304 > gpurun_out/ncu_vec.log 2>&1; echo "ncu vec rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -15 gpurun_out/t_gpu_all.log; cat gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err; cat gpurun_out/bench_reference.json; cat gpurun_out/vec_probe.log
