#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_vec.py tests/test_gpu_blas.py -x -q -m gpu -k "dot or gemv or vec" > gpurun_out/t_vec.log 2>&1; echo "vec tests rc=$?"
timeout 600 python bench.py --workload dot16m_212bit --no-e2e --no-cpu-baseline > gpurun_out/bench_dot16m_212bit.json 2> gpurun_out/bench_dot.err; echo "bench dot rc=$?"
tail -5 gpurun_out/t_vec.log | cut -c1-200; cut -c1-900 gpurun_out/bench_dot16m_212bit.json
