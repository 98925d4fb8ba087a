#!/bin/bash
# compute-sanitizer over the lone-query kernel (maxsim_rowm_kernel): memcheck + synccheck on the ragged 9-page case of every dtype,
# racecheck on smoke().  gpurun -- 'bash tools/gpu_sanit_rowm.sh'
set -u
mkdir -p gpurun_out
: > gpurun_out/sanitizer_rowm.txt
for tool in memcheck synccheck; do
  timeout -s KILL 400 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rows_as_m and 64-9" > gpurun_out/sanit_rowm_$tool.log 2>&1
  echo "== $tool rc=$?" >> gpurun_out/sanitizer_rowm.txt; grep -E "ERROR SUMMARY|passed|failed|error" gpurun_out/sanit_rowm_$tool.log | tail -4 >> gpurun_out/sanitizer_rowm.txt
done
timeout -s KILL 400 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/sanit_rowm_race.log 2>&1
echo "== racecheck smoke rc=$?" >> gpurun_out/sanitizer_rowm.txt; grep -E "RACECHECK SUMMARY|smoke|hazard" gpurun_out/sanit_rowm_race.log | tail -4 >> gpurun_out/sanitizer_rowm.txt
cat gpurun_out/sanitizer_rowm.txt
