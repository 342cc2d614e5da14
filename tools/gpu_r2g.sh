#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_fullprec.py -q -m gpu > gpurun_out/t_fullprec.log 2>&1; echo "fullprec rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --workload gemm1024_106bit --full-precision-inputs --no-e2e --no-sub --no-cpu-baseline > gpurun_out/b_c2_full.json 2> gpurun_out/b_c2_full.err; echo "bench c2 full rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -25 gpurun_out/t_fullprec.log | cut -c1-500; grep "^{" gpurun_out/b_c2_full.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d[\"ms_per_step\"], d.get(\"per_kernel_ms\"), d.get(\"verified_mismatches\"), d.get(\"worst_error_over_bound\"))"; tail -3 gpurun_out/b_c2_full.err
