#!/bin/bash
# round 2: new operations (norms, SpMV, conversions), full-precision tests, the GEMV / DOT tests (kernels_vec.cuh changed)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu > gpurun_out/t_ops.log 2>&1; echo "ops rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_fullprec.py -q -m gpu > gpurun_out/t_fullprec.log 2>&1; echo "fullprec rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_vec.py tests/test_gpu_level1.py -q -m gpu > gpurun_out/t_vec.log 2>&1; echo "vec rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --workload gemm1024_106bit --full-precision-inputs --no-e2e --no-sub --no-cpu-baseline > gpurun_out/b_c2_full.json 2> gpurun_out/b_c2_full.err; echo "bench c2 full rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -40 gpurun_out/t_ops.log | cut -c1-500; tail -12 gpurun_out/t_fullprec.log | cut -c1-400; tail -5 gpurun_out/t_vec.log | cut -c1-300
grep '^{' gpurun_out/b_c2_full.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in d if k.startswith('verif') or k.startswith('worst') or k.startswith('bitwise')})"; tail -3 gpurun_out/b_c2_full.err
