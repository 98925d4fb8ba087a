#!/bin/bash
# Round-2 GPU check.  Usage: gpurun -- 'bash tools/gpu_round2.sh [quick|full|prof]'
#   quick: parity tests (-x), smoke, bench on a 65536-page shard          full: all tests, smoke, the default bench
#   prof : ncu launch list of the bench + ncu --set full captures of the dominant kernels (feed to tools/ncu_summary.py)
set -u
mode=${1:-quick}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
if [ "$mode" = sanit ]; then
  # compute-sanitizer over the kernels that are new this round (fde_scan_umma, clamp / rerank_batch plumbing, gathered merge,
  # CUDA-graph replay, fp8) -- results go to gpurun_out/sanitizer_r02.txt
  : > gpurun_out/sanitizer_r02.txt
  for tool in memcheck synccheck; do
    timeout -s KILL 900 compute-sanitizer --tool $tool python -m pytest tests -m gpu -q -x -k "fde_scan_tensor_path and 17 or rerank_batch_per_query and bf16 or zero_pad_compat_rerank or sharded_search_pipeline_world1 or host_graph_replay and bf16 or fp8_matches_oracle and 13 or cta_pair_form_agrees and fp8 and 57" > gpurun_out/sanit_$tool.log 2>&1
    echo "== $tool rc=$?" >> gpurun_out/sanitizer_r02.txt; grep -E "ERROR SUMMARY|passed|failed|error" gpurun_out/sanit_$tool.log | tail -5 >> gpurun_out/sanitizer_r02.txt
  done
  timeout -s KILL 600 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/sanit_race.log 2>&1
  echo "== racecheck smoke rc=$?" >> gpurun_out/sanitizer_r02.txt; grep -E "RACECHECK SUMMARY|smoke" gpurun_out/sanit_race.log | tail -3 >> gpurun_out/sanitizer_r02.txt
  cat gpurun_out/sanitizer_r02.txt
  exit 0
fi
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
