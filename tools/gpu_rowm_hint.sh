#!/bin/bash
# A/B of the suspend-time hint on the mbarrier waits of maxsim_rowm_kernel.  gpurun -- 'bash tools/gpu_rowm_hint.sh'
set -u
mkdir -p gpurun_out
for v in 0 1 0 1; do
  B200MS_ROWM_HINT=$v timeout -s KILL 200 python tools/time_scan.py --dtypes bf16,int8,fp8,binary --pages 65536 --steps 20 2> gpurun_out/rowm_hint_$v.err | sed "s/^/{\"wait_hint\": $v} /" | tee -a gpurun_out/rowm_hint_ab.jsonl | cut -c1-600
done
