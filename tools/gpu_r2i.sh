#!/bin/bash
# round 2: scaling record with the flat layout: sharded mp_gemm at 8 / 4 GPUs, panel groups 1 / 2 / 4, per-panel packages for comparison
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { # name, ngpus, env..., then args
  local name=$1 n=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 5 --warmup 3 --no-sub --no-e2e --no-cpu-baseline "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?" >> gpurun_out/summary.txt
}
run f8_g2 8
MPRES_SHARD_GROUPS=1 run f8_g1 8
MPRES_SHARD_GROUPS=4 run f8_g4 8
MPRES_SHARD_GROUPS=8 run f8_g8 8
MPRES_PUSH_STREAMS=4 run f8_g2_p4 8
MPRES_SHARD_FLAT=0 run f8_pkg 8
run f4_g2 4
MPRES_SHARD_GROUPS=1 run f4_g1 4
MPRES_SHARD_GROUPS=4 run f4_g4 4
cat gpurun_out/summary.txt
for f in f8_g2 f8_g1 f8_g4 f8_g8 f8_g2_p4 f8_pkg f4_g2 f4_g1 f4_g4; do echo $f; grep '^{' gpurun_out/$f.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('per_kernel_ms'), d.get('verified_mismatches'))"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/$f.err | tail -3; done
