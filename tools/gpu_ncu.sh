#!/bin/bash
# launch list + ncu --set full captures; the .ncu-rep files are converted to CSV on the box and dropped
# (gpurun_out/ is capped at 64 MiB)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/*.ncu-rep
MPRES_BENCH_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> gpurun_out/summary.txt
MPRES_BENCH_PROFILER_RANGE=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none -k regex:'k_limb_umma|k_small_umma|k_norm_fast|k_align|k_minplus|k_norm_list|k_base_extend|k_ext_small|k_outer_info' -c 10 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/summary.txt
ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv 2>/dev/null
VEC_CFGS=0 timeout 900 ncu --set full --clock-control none -k regex:'k_mv_acc|k_mv_fin' -c 12 -o gpurun_out/prof_vec -f python tools/vec_probe.py 16 8192 8192 4194304 > gpurun_out/ncu_vec.log 2>&1; echo "ncu vec rc=$?" >> gpurun_out/summary.txt
ncu -i gpurun_out/prof_vec.ncu-rep --page raw --csv > gpurun_out/prof_vec_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep >> gpurun_out/summary.txt
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/summary.txt
