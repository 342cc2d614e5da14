#!/bin/bash
# first GPU check of the tcgen05 stage-2 kernel: parity of the three stage-2 kernels, then A/B bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_blas.py -x -q -m gpu -k "stage2" > gpurun_out/t_stage2.log 2>&1; echo "stage2 tests rc=$?" >> gpurun_out/summary.txt
for k in mma_sync umma_unstacked umma; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --stage2 $k > gpurun_out/bench_$k.json 2> gpurun_out/bench_$k.err; echo "bench $k rc=$?" >> gpurun_out/summary.txt
done
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "all gpu tests rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/t_stage2.log; cat gpurun_out/bench_*.json
