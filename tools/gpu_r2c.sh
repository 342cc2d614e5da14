#!/bin/bash
# round 2: scaling record on one 8-GPU box: sharded mp_gemm at 8 and 4 GPUs (B resident), the B-on-rank-0 convention at 8, DOT / GEMV at 8
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { # name, ngpus, args...
  local name=$1 n=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 5 --warmup 3 --no-sub "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?" >> gpurun_out/summary.txt
}
run s8_sharded 8 --no-e2e
run s4_sharded 4 --no-e2e
run s8_full 8 --no-e2e --bcast full
run s8_dot 8 --no-e2e --workload dot16m_212bit
run s8_gemv 8 --no-e2e --workload gemv16384_212bit
run s8_gemvt 8 --no-e2e --workload gemvt16384_212bit
run s8_e2e 8 --e2e-steps 2
cat gpurun_out/summary.txt
for f in s8_sharded s4_sharded s8_full s8_dot s8_gemv s8_gemvt s8_e2e; do grep '^{' gpurun_out/$f.json | tail -1 | cut -c1-160; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/$f.err | tail -4; done
