#!/bin/bash
# rows-as-M lone-query kernel: parity tests first, then A/B timing against the query-as-M kernels.  gpurun -- 'bash tools/gpu_rowm.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -1
timeout -s KILL 420 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rows_as_m or binary_bit_exact or ragged or config0 or int8_bit_exact or fp8_matches" 2>&1 | tail -30 > gpurun_out/pytest_rowm.txt
tail -12 gpurun_out/pytest_rowm.txt
for v in 0 1; do
  B200MS_ROWM=$v timeout -s KILL 300 python tools/time_scan.py --dtypes bf16,int8,fp8,binary --pages ${1:-65536} 2> gpurun_out/rowm_$v.err | sed "s/^/{\"rowm\": $v} /" | tee -a gpurun_out/rowm_ab.jsonl | cut -c1-600
  tail -2 gpurun_out/rowm_$v.err
done
