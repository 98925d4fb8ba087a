#!/bin/bash
# 8-GPU run (gpurun --gpus 8): bench at N=8 (equal shards, all legs), N=8 with speed-proportional placement (headline only),
# the sharded plugin under 32 concurrent callers, and the sharded two-stage store.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run() { timeout -s KILL "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
run 900 29521 bench.py --gpus 8 --steps ${STEPS:-60} --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$? lines=$(wc -l < gpurun_out/bench_n8.json)"; tail -4 gpurun_out/bench_n8.err
run 600 29522 bench.py --gpus 8 --steps ${STEPS:-60} --warmup 5 --placement speed --skip-legs > gpurun_out/bench_n8_speed.json 2> gpurun_out/bench_n8_speed.err; echo "bench n8 speed rc=$?"; tail -3 gpurun_out/bench_n8_speed.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_n8.json','gpurun_out/bench_n8_speed.json'):
    try:
        d=json.load(open(f))
        print(f, 'value %.4g e2e %.4g ms/step %.2f placement %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['placement']))
        print('  per_rank', d['per_rank']); print('  check', d['multi_gpu_check'])
        if d.get('config3_bq256'): print('  cfg3', d['config3_bq256'])
        if d.get('config4_two_stage'): print('  two_stage', {k:v for k,v in d['config4_two_stage'].items() if k in ('p50_ms','p95_ms','stage_ms_p50','recall','error')})
    except Exception as e: print(f, 'parse failed', e)
PY
run 600 29523 tools/bench_concurrency.py --sharded --clients 32 --rounds 15 --pages 65536 > gpurun_out/conc_n8.json 2> gpurun_out/conc_n8.err; echo "concurrency n8 rc=$?"; tail -c 2000 gpurun_out/conc_n8.json; tail -3 gpurun_out/conc_n8.err
run 600 29524 tools/bench_concurrency.py --sharded --clients 8 --rounds 15 --pages 65536 --fde-candidates 1000 > gpurun_out/conc_n8_two_stage.json 2> gpurun_out/conc_n8_two_stage.err; echo "two-stage concurrency n8 rc=$?"; tail -c 1500 gpurun_out/conc_n8_two_stage.json; tail -3 gpurun_out/conc_n8_two_stage.err
