#!/bin/bash
# A/B of the epilogue fast path of maxsim_rowm_kernel per corpus dtype.  gpurun -- 'bash tools/gpu_rowm_fast.sh'
set -u
mkdir -p gpurun_out
for v in 0 1 0 1; do
  B200MS_ROWM_FAST=$v timeout -s KILL 300 python tools/time_scan.py --dtypes bf16,int8,fp8,binary --pages 65536 2> gpurun_out/rowm_fast_$v.err | sed "s/^/{\"fast_path\": $v} /" | tee -a gpurun_out/rowm_fast_ab.jsonl | cut -c1-600
done
