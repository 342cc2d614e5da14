#!/bin/bash
# per-file durations of the GPU suite (files run concurrently: the oracle side of most tests is host work) and the L2 fetch-granularity experiment
mkdir -p gpurun_out/dur
rm -f gpurun_out/dur/*
pids=()
for f in tests/test_gpu_*.py; do
  b=$(basename $f .py)
  ( /usr/bin/time -f "$b wall %e s" timeout 1500 python -m pytest $f -q -m gpu --durations=12 > gpurun_out/dur/$b.log 2>&1; echo "$b rc=$?" >> gpurun_out/dur/summary.txt ) 2>> gpurun_out/dur/wall.txt &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
cat gpurun_out/dur/summary.txt gpurun_out/dur/wall.txt
for g in 0 32 64 128; do
  if [ $g = 0 ]; then unset MPRES_L2_FETCH; else export MPRES_L2_FETCH=$g; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-sub --no-e2e --no-cpu-baseline --no-verify > gpurun_out/dur/l2_$g.json 2> gpurun_out/dur/l2_$g.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/dur/l2_$g.json").read().strip().splitlines()[-1])
    print("L2_FETCH=$g", d["ms_per_step"], json.dumps(d.get("per_kernel_ms")))
except Exception as e:
    print("L2_FETCH=$g failed", e)
PY
done
