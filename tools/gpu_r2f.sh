#!/bin/bash
# round 2: sharded mp_gemm with the flat receive layout: parity on 2 GPUs, bench at 2 GPUs (flat / per-panel packages)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/t_sharded.log 2>&1; echo "sharded rc=$?" >> gpurun_out/summary.txt
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-sub --no-cpu-baseline > gpurun_out/b2_flat.json 2> gpurun_out/b2_flat.err; echo "bench 2 flat rc=$?" >> gpurun_out/summary.txt
MPRES_SHARD_FLAT=0 timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-sub --no-cpu-baseline > gpurun_out/b2_pkg.json 2> gpurun_out/b2_pkg.err; echo "bench 2 packages rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -15 gpurun_out/t_sharded.log | cut -c1-300
for f in b2_flat b2_pkg; do grep '^{' gpurun_out/$f.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('per_kernel_ms'), d.get('verified_mismatches'))"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/$f.err | tail -4; done
