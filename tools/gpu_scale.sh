#!/bin/bash
# scaling record: bench.py at N GPUs (N = first argument): lean broadcast without / with prefetch, full broadcast, DOT, GEMV
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > gpurun_out/scale_${N}_lean.json 2> gpurun_out/scale_${N}_lean.err; echo "N=$N lean rc=$?"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --prefetch > gpurun_out/scale_${N}_lean_prefetch.json 2> gpurun_out/scale_${N}_lean_prefetch.err; echo "N=$N lean prefetch rc=$?"
if [ "$2" != "short" ]; then
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --bcast full > gpurun_out/scale_${N}_full.json 2> gpurun_out/scale_${N}_full.err; echo "N=$N full rc=$?"
timeout 600 $TR bench.py --gpus $N --workload dot16m_212bit --no-e2e > gpurun_out/scale_${N}_dot.json 2> gpurun_out/scale_${N}_dot.err; echo "N=$N dot rc=$?"
timeout 600 $TR bench.py --gpus $N --workload gemv16384_212bit --no-e2e > gpurun_out/scale_${N}_gemv.json 2> gpurun_out/scale_${N}_gemv.err; echo "N=$N gemv rc=$?"
fi
for f in gpurun_out/scale_${N}_*.json; do echo $f; tail -1 $f | cut -c1-220; done; tail -3 gpurun_out/scale_${N}_lean.err
