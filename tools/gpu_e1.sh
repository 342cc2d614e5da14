#!/bin/bash
# quick A/B after a kernel change: parity of the small-base path, the default workload's per-kernel times, an ncu capture of the GEMV^T kernel
mkdir -p gpurun_out/e1; rm -f gpurun_out/e1/*
timeout 600 python -m pytest tests/test_gpu_small.py tests/test_gpu_scale_parity.py tests/test_gpu_fullprec.py -q -m gpu -x > gpurun_out/e1/tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/e1/tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-sub --no-e2e --no-cpu-baseline > gpurun_out/e1/bench.json 2> gpurun_out/e1/bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/e1/bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d.get("verified_mismatches"), json.dumps(d.get("per_kernel_ms")))
PY
if [ -n "$1" ]; then
MPRES_BENCH_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"$1" -c 3 -o gpurun_out/e1/prof -f python bench.py --workload $2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub --no-verify > gpurun_out/e1/ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/e1/prof.ncu-rep --page raw --csv > gpurun_out/e1/prof_raw.csv 2>/dev/null
ncu -i gpurun_out/e1/prof.ncu-rep --page source --csv > gpurun_out/e1/prof_source.csv 2>/dev/null
rm -f gpurun_out/e1/prof.ncu-rep
fi
