#!/bin/bash
# per-file durations of the GPU suite (files run concurrently: the oracle side of most tests is host work)
mkdir -p gpurun_out/dur
rm -f gpurun_out/dur/*
pids=()
for f in tests/test_gpu_*.py; do
  b=$(basename $f .py)
  ( t0=$(date +%s); timeout 1500 python -m pytest $f -q -m gpu --durations=12 > gpurun_out/dur/$b.log 2>&1; echo "$b rc=$? wall $(( $(date +%s) - t0 )) s" >> gpurun_out/dur/summary.txt ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
cat gpurun_out/dur/summary.txt
