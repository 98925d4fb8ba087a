#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 900 python tools/bench_two_stage.py --pages 65536 > gpurun_out/two_stage.json 2> gpurun_out/two_stage.err; echo "two_stage rc=$?"; cat gpurun_out/two_stage.json; tail -3 gpurun_out/two_stage.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
