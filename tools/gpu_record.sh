#!/bin/bash
# round-2 record on one B200: smoke, the GPU suite (files side by side: most of their time is the host-side oracle), the default bench line
# (sub_results, e2e, cpu_baseline), the reference arm, the ncu launch list and the ncu --set full capture of the same bench command
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
bash tools/gpu_durations.sh >> gpurun_out/summary.txt 2>&1
( time timeout 1200 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?" >> gpurun_out/summary.txt
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?" >> gpurun_out/summary.txt
MPRES_BENCH_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-sub --no-verify > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> gpurun_out/summary.txt
MPRES_BENCH_PROFILER_RANGE=1 timeout 1200 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'k_small_umma|k_align_small|k_ext_small|k_norm_fast|k_mp_gather|k_outer' -c 12 -o gpurun_out/prof_final -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub --no-verify > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/summary.txt
ncu -i gpurun_out/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench_default.json; tail -4 gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_reference.json
