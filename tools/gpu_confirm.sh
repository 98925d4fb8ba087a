#!/bin/bash
# last confirmation of HEAD on a B200: all GPU tests, smoke, the default bench (no profiler).  gpurun -- 'bash tools/gpu_confirm.sh'
set -u
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_head.txt; tail -2 gpurun_out/pytest_gpu_head.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_n1_head.json 2> gpurun_out/bench_head.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench_n1_head.json)"; tail -2 gpurun_out/bench_head.err
