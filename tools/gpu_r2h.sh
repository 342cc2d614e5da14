#!/bin/bash
# round 2: single-rounding stage 3 with the binary epilogue and sliced significands (N = 8 ... 32), GEMM tests, p-bit benches, int8 peak
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_fullprec.py -q -m gpu > gpurun_out/t_fullprec.log 2>&1; echo "fullprec rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_blas.py tests/test_gpu_small.py tests/test_gpu_fuzz.py -q -m gpu > gpurun_out/t_blas.log 2>&1; echo "blas rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --workload gemm1024_106bit --full-precision-inputs --no-e2e --no-sub --no-cpu-baseline > gpurun_out/b_c2_full.json 2> gpurun_out/b_c2_full.err; echo "bench c2 full rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --full-precision-inputs --no-e2e --no-sub --no-cpu-baseline > gpurun_out/b_c3_full.json 2> gpurun_out/b_c3_full.err; echo "bench c3 full rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --no-e2e --no-sub --no-cpu-baseline > gpurun_out/b_c3.json 2> gpurun_out/b_c3.err; echo "bench c3 rc=$?" >> gpurun_out/summary.txt
timeout 120 tools/int8_peak > gpurun_out/int8_peak.json 2> gpurun_out/int8_peak.err; echo "int8 peak rc=$?" >> gpurun_out/summary.txt
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv >> gpurun_out/int8_peak.err
cat gpurun_out/summary.txt; tail -30 gpurun_out/t_fullprec.log | cut -c1-500; tail -8 gpurun_out/t_blas.log | cut -c1-300
for f in b_c2_full b_c3_full b_c3; do grep "^{" gpurun_out/$f.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d.get('per_kernel_ms'), d.get('verified_entries'), d.get('verified_mismatches'), d.get('worst_error_over_bound'), d.get('fallback_elements_last_step'), d.get('small_base_moduli'))"; tail -3 gpurun_out/$f.err; done
cat gpurun_out/int8_peak.json gpurun_out/int8_peak.err
