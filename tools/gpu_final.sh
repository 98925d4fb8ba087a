#!/bin/bash
# Final round-2 GPU pass: all GPU tests, smoke, the default bench, the ncu launch list, ncu --set full captures of the lone-query
# kernel (bf16 / int8 / sign bits).  gpurun --timeout 1500 -- 'bash tools/gpu_final.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -1
timeout -s KILL 600 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -30 > gpurun_out/pytest_gpu_final2.txt; tail -3 gpurun_out/pytest_gpu_final2.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_n1_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench_n1_final2.json)"; tail -2 gpurun_out/bench_final2.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'maxsim|topk|pack_rows|chunk_page|merge|b1_query|rowm_query|fde_' -c 400 --csv --log-file gpurun_out/launches_bench2.csv python bench.py --pages 65536 --steps 3 --warmup 3 --no-cpu-baseline --sweep-pages 16384 --fde-pages 32768 --topic-pages 16384 --latency-queries 8 > gpurun_out/bench_under_ncu2.log 2>&1; echo "ncu launches rc=$?"
for dt in bf16 int8 binary; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:maxsim_rowm_kernel -s 1 -c 1 -o gpurun_out/prof_rowm_$dt -f python tools/profile_kernels.py --only $dt --bqs 1 --pages 32768 > gpurun_out/ncu_rowm_$dt.log 2>&1; echo "ncu rowm $dt rc=$?"
done
