#!/bin/bash
set -u
for v in 0 1 0 1; do
  B200MS_SPLIT4=$v timeout 900 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/b_$v.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/b_$v.json'))
print('split4=$v value %.4g tensor %.1f | hbm %.0f GB/s step %.2f score %.2f | clocks %s' % (d['value'], d['roofline']['achieved'], d['hbm_regime']['achieved'], d['hbm_regime']['step_ms'], d['hbm_regime']['score_ms'], d['clocks']['sm_mhz']))
PY
done
