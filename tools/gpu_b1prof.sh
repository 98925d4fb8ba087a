#!/bin/bash
set -u
B200MS_B1_TENSOR=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_b1_umma -s 5 -c 1 -o gpurun_out/prof_b1_umma -f python tools/profile_kernels.py --binary --pages 32768 > gpurun_out/ncu_b1.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_b1.log
