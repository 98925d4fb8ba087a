#!/bin/bash
# Round-2 GPU check: parity tests, smoke, a quick bench on a small shard.  Usage: gpurun -- 'bash tools/gpu_round2.sh [quick|full]'
set -u
mode=${1:-quick}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
timeout -s KILL 1800 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt; tail -12 gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
if [ "$mode" = quick ]; then
  timeout 900 python bench.py --pages 65536 --steps 20 --sweep-pages 16384 --fde-pages 32768 --topic-pages 16384 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench quick rc=$? lines=$(wc -l < gpurun_out/bench_quick.json)"; tail -3 gpurun_out/bench_quick.err
else
  timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench.json)"; tail -3 gpurun_out/bench.err
fi
