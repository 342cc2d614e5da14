#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/ref_gpu_compare.py > gpurun_out/ref_gpu_compare.jsonl 2> gpurun_out/ref_gpu_compare.err; echo "ref compare rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --stage3 3 > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; echo "bench fused rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench rc=$?"
cat gpurun_out/ref_gpu_compare.jsonl; tail -3 gpurun_out/ref_gpu_compare.err; cut -c1-200 gpurun_out/bench_fused.json; cut -c1-200 gpurun_out/bench_small.json
