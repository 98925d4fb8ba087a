#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$? lines=$(wc -l < gpurun_out/bench_n2.json)"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print('N=2 value %.4g e2e %.4g tensor %.1f TF frac %.3f | launches %d | clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['gpu_launches'], d['clocks']))
PY
