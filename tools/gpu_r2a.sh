#!/bin/bash
# round 2, first record: smoke, the new parity tests, the whole GPU suite, the default bench line (with sub_results)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_scale_parity.py -x -q -m gpu > gpurun_out/t_scale.log 2>&1; echo "scale parity rc=$?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu --deselect tests/test_gpu_scale_parity.py > gpurun_out/t_gpu_all.log 2>&1; echo "all gpu tests rc=$?" >> gpurun_out/summary.txt
( time timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log; tail -15 gpurun_out/t_scale.log | cut -c1-300; tail -15 gpurun_out/t_gpu_all.log | cut -c1-300; cut -c1-1500 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
