#!/bin/bash
# Round-2 GPU check.  Usage: gpurun -- 'bash tools/gpu_round2.sh [quick|full|prof]'
#   quick: parity tests (-x), smoke, bench on a 65536-page shard          full: all tests, smoke, the default bench
#   prof : ncu launch list of the bench + ncu --set full captures of the dominant kernels (feed to tools/ncu_summary.py)
set -u
mode=${1:-quick}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
if [ "$mode" != prof ]; then
  x=""; [ "$mode" = quick ] && x="-x"
  timeout -s KILL 1800 python -m pytest tests -m gpu -q $x --durations=8 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt; tail -12 gpurun_out/pytest_gpu.txt
  timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
fi
if [ "$mode" = quick ]; then
  timeout 900 python bench.py --pages 65536 --steps 20 --sweep-pages 16384 --fde-pages 32768 --topic-pages 16384 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench quick rc=$? lines=$(wc -l < gpurun_out/bench_quick.json)"; tail -3 gpurun_out/bench_quick.err
elif [ "$mode" = full ]; then
  timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench.json)"; tail -3 gpurun_out/bench.err
fi
if [ "$mode" = prof ] || [ "$mode" = full ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'maxsim|topk|pack_rows|chunk_page|merge|b1_query|fde_' -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --pages 65536 --steps 3 --warmup 3 --no-cpu-baseline --sweep-pages 16384 --fde-pages 32768 --topic-pages 16384 --latency-queries 8 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_umma_pair_kernel -s 2 -c 1 -o gpurun_out/prof_umma_pair -f python tools/profile_kernels.py > gpurun_out/ncu_pair.log 2>&1; echo "ncu pair rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'maxsim_umma_kernel' -s 2 -c 1 -o gpurun_out/prof_umma_nm1 -f python tools/profile_kernels.py > gpurun_out/ncu_nm1.log 2>&1; echo "ncu nm1 rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_umma_pair_kernel -s 2 -c 1 -o gpurun_out/prof_umma_pair_int8 -f python tools/profile_kernels.py --only int8 > gpurun_out/ncu_pair_int8.log 2>&1; echo "ncu pair int8 rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_umma_pair_kernel -s 2 -c 1 -o gpurun_out/prof_umma_pair_fp8 -f python tools/profile_kernels.py --only fp8 > gpurun_out/ncu_pair_fp8.log 2>&1; echo "ncu pair fp8 rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fde_scan_umma_kernel -s 2 -c 1 -o gpurun_out/prof_fde_scan -f python tools/profile_kernels.py --only fde > gpurun_out/ncu_fde.log 2>&1; echo "ncu fde rc=$?"
fi
