#!/bin/bash
# round 2, second record (2 GPUs): sharded parity test, the tests that failed in the first record, bench at 2 GPUs (sharded and B-broadcast conventions)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/t_sharded.log 2>&1; echo "sharded rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_small.py tests/test_gpu_scale_parity.py -q -m gpu > gpurun_out/t_fixed.log 2>&1; echo "fixed tests rc=$?" >> gpurun_out/summary.txt
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-sub > gpurun_out/b2_sharded.json 2> gpurun_out/b2_sharded.err; echo "bench 2 sharded rc=$?" >> gpurun_out/summary.txt
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-sub --bcast full > gpurun_out/b2_full.json 2> gpurun_out/b2_full.err; echo "bench 2 bcast rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -15 gpurun_out/t_sharded.log | cut -c1-300; tail -8 gpurun_out/t_fixed.log | cut -c1-300
for f in b2_sharded b2_full; do tail -1 gpurun_out/$f.json | cut -c1-1200; tail -5 gpurun_out/$f.err; done
