#!/bin/bash
# default bench line (e2e + cpu_baseline), reference arm, config-2 and precision-sweep lines
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
( time timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?" >> gpurun_out/summary.txt
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?" >> gpurun_out/summary.txt
for w in gemm1024_106bit gemm2048_106bit gemm2048_212bit gemm2048_318bit gemm2048_424bit gemm2048_530bit gemm2048_636bit gemm2048_742bit gemm2048_848bit; do
  timeout 600 python bench.py --workload $w --no-e2e --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?" >> gpurun_out/summary.txt
done
timeout 600 python bench.py --full-precision-inputs --workload gemm1024_106bit --no-e2e --no-cpu-baseline > gpurun_out/bench_gemm1024_fullprec.json 2> gpurun_out/bench_gemm1024_fullprec.err; echo "bench fullprec rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err; cat gpurun_out/bench_reference.json
