#!/bin/bash
# 8-GPU run (gpurun --gpus 8): bench at N=8 (equal shards, all legs), N=8 with speed-proportional placement (headline only),
# the sharded plugin under 32 concurrent callers, and the sharded two-stage store.  Every step is time-boxed and carries a
# faulthandler watchdog (a hung collective dumps its stacks and exits: 8 GPUs idling on a lease is the expensive failure).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run() { B200MS_WATCHDOG_S="$1" timeout -s KILL "$2" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port "$3" "${@:4}"; }
run 190 205 29521 bench.py --gpus 8 --steps ${STEPS:-40} --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$? lines=$(wc -l < gpurun_out/bench_n8.json)"; grep -v "^\[W" gpurun_out/bench_n8.err | grep -i "error\|assert\|File \"/" | tail -6
run 85 95 29523 tools/bench_concurrency.py --sharded --clients 32 --rounds 10 --pages 65536 > gpurun_out/conc_n8.json 2> gpurun_out/conc_n8.err; echo "concurrency n8 rc=$?"; tail -c 2000 gpurun_out/conc_n8.json; grep -v "^\[W\|^$" gpurun_out/conc_n8.err | tail -4
run 95 105 29522 bench.py --gpus 8 --steps ${STEPS:-40} --warmup 5 --placement speed --skip-legs > gpurun_out/bench_n8_speed.json 2> gpurun_out/bench_n8_speed.err; echo "bench n8 speed rc=$?"; grep -v "^\[W" gpurun_out/bench_n8_speed.err | grep -i "error\|assert" | tail -3
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
