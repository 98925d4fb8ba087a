#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench.json)"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value %.4g e2e %.4g tensor %.1f TF frac %.3f | hbm %.0f GB/s frac %.3f step %.2f score %.2f | launches %d | cpu %.3g | clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['hbm_regime']['achieved'], d['hbm_regime']['frac'], d['hbm_regime']['step_ms'], d['hbm_regime']['score_ms'], d['gpu_launches'], d['cpu_baseline']['value'], d['clocks']))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'maxsim|topk|pack_rows|chunk_page|merge|b1_query' -c 300 --csv --log-file gpurun_out/launches_bench.csv python bench.py --pages 65536 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 600 python tools/bench_two_stage.py --pages 65536 2>/dev/null > gpurun_out/two_stage.json; cut -c1-420 gpurun_out/two_stage.json
