#!/bin/bash
set -u
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_q.json'))
print('value %.4g e2e %.4g tensor %.1f TF frac %.3f ms/step %.2f score %.2f | hbm %.0f GB/s frac %.3f step %.2f score %.2f | clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['ms_per_step'], d['roofline']['score_ms_per_step'], d['hbm_regime']['achieved'], d['hbm_regime']['frac'], d['hbm_regime']['step_ms'], d['hbm_regime']['score_ms'], d['clocks']))
PY
timeout 600 python tools/bench_two_stage.py --pages 65536 2>/dev/null | cut -c1-420
