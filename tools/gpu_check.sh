#!/bin/bash
# last check of the round's final library: smoke, the GPU suite (files side by side), the default bench line
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 100 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
bash tools/gpu_durations.sh >> gpurun_out/summary.txt 2>&1
timeout 100 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -1 gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench_default.json
