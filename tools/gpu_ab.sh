#!/bin/bash
# A/B of the one-CTA scorer's epilogue variants (tools/build_variants.py) + the full GPU test suite.  gpurun -- 'bash tools/gpu_ab.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -1
timeout -s KILL 600 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -25 > gpurun_out/pytest_gpu_ab.txt; tail -4 gpurun_out/pytest_gpu_ab.txt
timeout -s KILL 500 python tools/time_scan.py --variants "${1:-base,exact,one,exact_one,exact_one_spin,base}" > gpurun_out/ab_scan.jsonl 2> gpurun_out/ab_scan.err
cat gpurun_out/ab_scan.jsonl | cut -c1-400; tail -3 gpurun_out/ab_scan.err
