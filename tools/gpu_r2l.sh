#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_host_gemm.py tests/test_compat_shim.py -q -m gpu > gpurun_out/t_host.log 2>&1; echo "host gemm rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --no-sub --no-cpu-baseline --e2e-steps 3 > gpurun_out/b_e2e_lean.json 2> gpurun_out/b_e2e_lean.err; echo "bench e2e lean rc=$?" >> gpurun_out/summary.txt
MPRES_HOST_LEAN=0 timeout 600 python bench.py --no-sub --no-cpu-baseline --e2e-steps 3 > gpurun_out/b_e2e_full.json 2> gpurun_out/b_e2e_full.err; echo "bench e2e full rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -25 gpurun_out/t_host.log | cut -c1-400
for f in b_e2e_lean b_e2e_full; do grep "^{" gpurun_out/$f.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d.get('e2e'))"; tail -3 gpurun_out/$f.err; done
