#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none -k regex:'maxsim_umma|maxsim_b1' -s 8 -c 8 -o gpurun_out/prof_i8_b1 -f python tools/profile_kernels.py --int8 --binary --reps 1 > gpurun_out/ncu_i8b1.log 2>&1; echo "ncu rc=$?"; tail -8 gpurun_out/ncu_i8b1.log
