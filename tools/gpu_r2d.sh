#!/bin/bash
# round 2: full-precision (binary rounding) tests, the GEMM tests touched by the new outer-info kernels, config-2 bench with p-bit inputs
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_fullprec.py -x -q -m gpu > gpurun_out/t_fullprec.log 2>&1; echo "fullprec rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_blas.py tests/test_gpu_fuzz.py tests/test_gpu_small.py -q -m gpu > gpurun_out/t_blas.log 2>&1; echo "blas rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload gemm1024_106bit --full-precision-inputs --no-e2e --no-sub > gpurun_out/b_c2_full.json 2> gpurun_out/b_c2_full.err; echo "bench c2 full rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --no-e2e --no-sub --no-cpu-baseline > gpurun_out/b_c3.json 2> gpurun_out/b_c3.err; echo "bench c3 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -25 gpurun_out/t_fullprec.log | cut -c1-400; tail -8 gpurun_out/t_blas.log | cut -c1-300
for f in b_c2_full b_c3; do grep '^{' gpurun_out/$f.json | tail -1 | cut -c1-300; tail -3 gpurun_out/$f.err; done
