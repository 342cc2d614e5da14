#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_small.py -x -q -m gpu > gpurun_out/t_small.log 2>&1; echo "small tests rc=$?" >> gpurun_out/summary.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "all gpu tests rc=$?" >> gpurun_out/summary.txt
for v in small small_t128; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --stage2 $v > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?" >> gpurun_out/summary.txt
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --stage3 3 > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; echo "bench fused rc=$?" >> gpurun_out/summary.txt
timeout 1200 python tools/ref_gpu_compare.py > gpurun_out/ref_gpu_compare.jsonl 2> gpurun_out/ref_gpu_compare.err; echo "ref compare rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -30 gpurun_out/t_small.log | cut -c1-200; tail -6 gpurun_out/t_gpu_all.log | cut -c1-200; for f in small small_t128 fused; do cut -c1-200 gpurun_out/bench_$f.json; done; cat gpurun_out/ref_gpu_compare.jsonl; tail -3 gpurun_out/ref_gpu_compare.err
