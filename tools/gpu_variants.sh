#!/bin/bash
# A/B of library variants (tools/build_variant.sh): default workload, device-timed, sampled verification on
# usage: tools/gpu_variants.sh "name[:bench args]" ...     ("base" = the regular library)
mkdir -p gpurun_out/var
for spec in "$@"; do
  name=${spec%%:*}; extra=""; [ "$spec" != "$name" ] && extra=${spec#*:}
  if [ "$name" = base ]; then unset MPRES_B200_LIB; else export MPRES_B200_LIB=$PWD/mpres-blas_b200/libmpres_b200_$name.so; fi
  tag=$(echo "$spec" | tr ' :-' '___')
  timeout 300 python bench.py --steps 10 --warmup 3 --no-sub --no-e2e --no-cpu-baseline $extra > gpurun_out/var/$tag.json 2> gpurun_out/var/$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/var/$tag.json").read().strip().splitlines()[-1])
    print("%-22s %.4f ms  mismatches %s  %s" % ("$spec", d["ms_per_step"], d.get("verified_mismatches"), json.dumps(d.get("per_kernel_ms"))))
except Exception as e:
    print("$spec failed", e)
PY
done
