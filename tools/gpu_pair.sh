#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_gpu_parity.py -k "pair" -x -q 2>&1 | tail -15
rc=${PIPESTATUS[0]}
echo "pair tests rc=$rc"
if [ "$rc" != "0" ]; then
  timeout -s KILL 200 python tools/pair_debug.py 2>&1 | tail -60
  exit 0
fi
for p in 0 1 0 1; do
  echo "== profile_kernels PAIR=$p"; B200MS_PAIR_CTA=$p timeout -s KILL 400 python tools/profile_kernels.py --int8 2>&1 | grep "bq=32"
done
for p in 1 0; do
  B200MS_PAIR_CTA=$p timeout -s KILL 900 python bench.py --no-cpu-baseline > gpurun_out/bench_pair$p.json 2> gpurun_out/bench_pair$p.err; echo "bench PAIR=$p rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_pair$p.json'))
print('PAIR=$p value %.4g e2e %.4g tensor %.1f TF frac %.3f | hbm %.0f GB/s | launches %d | clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['hbm_regime']['achieved'], d['gpu_launches'], d['clocks']))
PY
done
B200MS_PAIR_CTA=1 timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_umma_pair -s 2 -c 1 -o gpurun_out/prof_umma_pair -f python tools/profile_kernels.py > gpurun_out/ncu_pair.log 2>&1; echo "ncu pair rc=$?"
