#!/bin/bash
# source-level ncu captures (one kernel each; the reports are converted to CSV on the box and dropped)
# usage: tools/gpu_ncu_src.sh [kernel-name-regex ...]   (default: k_norm_fast k_align_small)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/*.ncu-rep
KERNELS="${@:-k_norm_fast k_align_small}"
for kn in $KERNELS; do
MPRES_BENCH_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$kn -c 1 -o gpurun_out/src_$kn -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_src_$kn.log 2>&1; echo "ncu $kn rc=$?" >> gpurun_out/summary.txt
ncu -i gpurun_out/src_$kn.ncu-rep --page source --csv > gpurun_out/src_$kn.csv 2>/dev/null; ncu -i gpurun_out/src_$kn.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src2_$kn.csv 2>/dev/null
ncu -i gpurun_out/src_$kn.ncu-rep --page raw --csv > gpurun_out/raw_$kn.csv 2>/dev/null
ls -la gpurun_out/src_$kn.ncu-rep >> gpurun_out/summary.txt
rm -f gpurun_out/src_$kn.ncu-rep
done
cat gpurun_out/summary.txt; ls -la gpurun_out/src*.csv
