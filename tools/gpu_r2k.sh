#!/bin/bash
# p-bit config 3: stage-2 tile shapes / pair passes
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout 300 python bench.py --full-precision-inputs --no-e2e --no-sub --no-cpu-baseline --no-verify --steps 3 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pf_default
run pf_t128 --stage2 small_t128
MPRES_SLICE_PASSES=1 run pf_t128_passes --stage2 small_t128
cat gpurun_out/summary.txt
for f in pf_default pf_t128 pf_t128_passes; do grep "^{" gpurun_out/$f.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d.get('per_kernel_ms'))"; tail -2 gpurun_out/$f.err; done
