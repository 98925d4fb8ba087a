#!/bin/bash
# Full validation + final measurements for the round: tests, smoke, bench (+reference arm), ncu launch list, ncu captures.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value %.4g e2e %.4g tensor %.1f TF frac %.3f | hbm %.0f GB/s frac %.3f step %.2f score %.2f | clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['hbm_regime']['achieved'], d['hbm_regime']['frac'], d['hbm_regime']['step_ms'], d['hbm_regime']['score_ms'], d['clocks']))
PY
timeout 600 python tools/profile_kernels.py --int8 --binary > gpurun_out/profile_plain.log 2>&1; tail -6 gpurun_out/profile_plain.log
timeout 600 python tools/bench_two_stage.py --pages 65536 > gpurun_out/two_stage.json 2>/dev/null; cat gpurun_out/two_stage.json | cut -c1-700
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'maxsim|topk|pack_rows|chunk_page|merge|b1_query' -c 200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --pages 65536 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_umma -s 4 -c 2 -o gpurun_out/prof_umma_nm4 -f python tools/profile_kernels.py > gpurun_out/ncu_nm4.log 2>&1; echo "ncu nm4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_umma -s 10 -c 2 -o gpurun_out/prof_umma_nm1 -f python tools/profile_kernels.py > gpurun_out/ncu_nm1.log 2>&1; echo "ncu nm1 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'maxsim_b1_kernel|topk_kernel|fde_scan' -c 6 -o gpurun_out/prof_misc -f python tools/bench_two_stage.py --pages 16384 --queries 8 > gpurun_out/ncu_misc.log 2>&1; echo "ncu misc rc=$?"
