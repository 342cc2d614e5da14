#!/bin/bash
# 2-GPU: row-sharded GEMM bench (lean and full broadcast), segment-sharded DOT, row-sharded GEMV; 1-GPU A/B of two magnification rounds
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench2_lean.json 2> gpurun_out/bench2_lean.err; echo "bench 2gpu lean rc=$?" >> gpurun_out/summary.txt
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --bcast full > gpurun_out/bench2_full.json 2> gpurun_out/bench2_full.err; echo "bench 2gpu full rc=$?" >> gpurun_out/summary.txt
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench2_e2e.json 2> gpurun_out/bench2_e2e.err; echo "bench 2gpu e2e rc=$?" >> gpurun_out/summary.txt
timeout 600 $TR bench.py --gpus 2 --workload dot16m_212bit > gpurun_out/bench2_dot.json 2> gpurun_out/bench2_dot.err; echo "bench 2gpu dot rc=$?" >> gpurun_out/summary.txt
timeout 600 $TR bench.py --gpus 2 --workload gemv16384_212bit > gpurun_out/bench2_gemv.json 2> gpurun_out/bench2_gemv.err; echo "bench 2gpu gemv rc=$?" >> gpurun_out/summary.txt
timeout 300 $TR bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > gpurun_out/bench2_ref.json 2> gpurun_out/bench2_ref.err; echo "bench 2gpu ref rc=$?" >> gpurun_out/summary.txt
MPRES_B200_LIB=$PWD/build_ab/libmpres_b200_rounds2.so timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_rounds2.json 2> gpurun_out/bench_rounds2.err; echo "bench rounds2 rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_rounds1.json 2> gpurun_out/bench_rounds1.err; echo "bench rounds1 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; for f in gpurun_out/bench2_*.json gpurun_out/bench_rounds*.json; do echo $f; cut -c1-230 $f; done; tail -5 gpurun_out/bench2_lean.err
