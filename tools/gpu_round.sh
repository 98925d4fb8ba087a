#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for tool in memcheck synccheck; do
  timeout -s KILL 600 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -q -k "pair_form_agrees and 57" > gpurun_out/sanitizer_pair_$tool.txt 2>&1
  echo "sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_pair_$tool.txt | tail -3
done
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench.json)"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value %.4g e2e %.4g tensor %.1f TF frac %.3f | hbm %.0f GB/s frac %.3f | launches %d | torch_gpu bf16 %.3g | cpu %.3g | clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['hbm_regime']['achieved'], d['hbm_regime']['frac'], d['gpu_launches'], d['torch_gpu_reference_formulation']['bf16']['value'], d['cpu_baseline']['value'], d['clocks']))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'maxsim|topk|pack_rows|chunk_page|merge|b1_query' -c 300 --csv --log-file gpurun_out/launches_bench.csv python bench.py --pages 65536 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
# ncu --set full of the two dominant kernels (feed the .ncu-rep files to tools/ncu_summary.py here afterwards):
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_umma_pair_kernel -s 2 -c 1 -o gpurun_out/prof_umma_pair -f python tools/profile_kernels.py > gpurun_out/ncu_pair.log 2>&1; echo "ncu pair rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'maxsim_umma_kernel' -s 2 -c 1 -o gpurun_out/prof_umma_nm1 -f python tools/profile_kernels.py > gpurun_out/ncu_nm1.log 2>&1; echo "ncu nm1 rc=$?"
