#!/bin/bash
# round-2 scaling record: the default invocation (what the driver runs) at N GPUs
N=$1
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/scale_default_$N.json 2> gpurun_out/scale_default_$N.err
echo "scale default N=$N rc=$?"
grep '^{' gpurun_out/scale_default_$N.json | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['n_gpus'], d['ms_per_step'], d['value'], d.get('e2e'), d.get('verified_mismatches'), d.get('clocks'))
for s in d.get('sub_results', []): print(' ', s.get('name'), s.get('ms_per_step'), s.get('value'), s.get('error'))
"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/scale_default_$N.err | tail -6
