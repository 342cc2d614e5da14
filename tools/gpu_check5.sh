#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_vec.py -x -q > gpurun_out/t_vec.log 2>&1; echo "vec tests rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/vec_probe.py 16 16384 16384 > gpurun_out/vec_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/summary.txt
VEC_CFGS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_mv_acc' -c 12 -o gpurun_out/prof_vec -f python tools/vec_probe.py 16 8192 8192 4194304 > gpurun_out/ncu_vec.log 2>&1; echo "ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -30 gpurun_out/t_vec.log; cat gpurun_out/vec_probe.log
