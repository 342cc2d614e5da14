#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_fullprec.py -q -m gpu > gpurun_out/t_fullprec.log 2>&1; echo "fullprec rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "div or conjugate" > gpurun_out/t_ops.log 2>&1; echo "ops(div,cg) rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --full-precision-inputs --no-e2e --no-sub --no-cpu-baseline > gpurun_out/b_c3_full.json 2> gpurun_out/b_c3_full.err; echo "bench c3 full rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --workload gemm2048_212bit --full-precision-inputs --no-e2e --no-sub --no-cpu-baseline > gpurun_out/b_c5_212_full.json 2> gpurun_out/b_c5_212_full.err; echo "bench 2048/212 full rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --workload gemm1024_106bit --full-precision-inputs --no-e2e --no-sub --no-cpu-baseline > gpurun_out/b_c2_full.json 2> gpurun_out/b_c2_full.err; echo "bench c2 full rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -30 gpurun_out/t_fullprec.log | cut -c1-500; tail -40 gpurun_out/t_ops.log | cut -c1-400
for f in b_c3_full b_c5_212_full b_c2_full; do grep "^{" gpurun_out/$f.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d.get('per_kernel_ms'), d.get('verified_entries'), d.get('verified_mismatches'), d.get('worst_error_over_bound'), d.get('fallback_elements_last_step'), d.get('small_base_moduli'))"; tail -3 gpurun_out/$f.err; done
