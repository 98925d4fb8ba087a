#!/bin/bash
# ncu --set full captures of maxsim_rowm_kernel (sign bits and int8), one warm launch each.  gpurun -- 'bash tools/gpu_prof_rowm.sh'
set -u
mkdir -p gpurun_out
for dt in ${1:-binary int8}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:maxsim_rowm_kernel -s 1 -c 1 -o gpurun_out/prof_rowm_$dt -f python tools/profile_kernels.py --only $dt --bqs 1 --pages 16384 > gpurun_out/ncu_rowm_$dt.log 2>&1; echo "ncu rowm $dt rc=$?"
  tail -3 gpurun_out/ncu_rowm_$dt.log
done
