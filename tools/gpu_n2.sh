#!/bin/bash
# 2-GPU check (gpurun --gpus 2): NCCL tests (C-ABI communicator, pipelined search, sharded stores) + bench at N=2.
# Every step is time-boxed and every worker carries a faulthandler watchdog: a hung collective dumps its stacks and exits.
set -u
mkdir -p gpurun_out
mode=${1:-full}
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout -s KILL 420 python -m pytest tests/test_gpu_sharded_nccl.py -m gpu -q -x 2>&1 | tail -80 > gpurun_out/pytest_nccl.txt; tail -30 gpurun_out/pytest_nccl.txt
B200MS_WATCHDOG_S=150 timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_concurrency.py --sharded --clients 16 --rounds 10 --pages 32768 > gpurun_out/conc_n2.json 2> gpurun_out/conc_n2.err; echo "concurrency n2 rc=$?"; tail -c 1500 gpurun_out/conc_n2.json; grep -v "^\[W\|^$" gpurun_out/conc_n2.err | tail -25
[ "$mode" = tests ] && exit 0
pages=524288; [ "$mode" = quick ] && pages=65536
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps ${STEPS:-40} --warmup 5 --pages $pages > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$? lines=$(wc -l < gpurun_out/bench_n2.json)"; grep -v "^\[W" gpurun_out/bench_n2.err | grep -i "error\|assert" | tail -5
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_n2.json'))
    print('N=2 value %.4g e2e %.4g ms/step %.2f | per_rank %s | check %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['per_rank'], d['multi_gpu_check']))
    print('cfg3', d['config3_bq256']); print('two_stage', {k:v for k,v in (d['config4_two_stage'] or {}).items() if k in ('p50_ms','p95_ms','stage_ms_p50','recall','error')})
except Exception as e: print('parse failed', e)
PY
