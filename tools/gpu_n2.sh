#!/bin/bash
# 2-GPU check (gpurun --gpus 2): NCCL tests (C-ABI communicator, pipelined search, sharded stores) + bench at N=2.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout -s KILL 1200 python -m pytest tests/test_gpu_sharded_nccl.py -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_nccl.txt; tail -6 gpurun_out/pytest_nccl.txt
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps ${STEPS:-40} --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$? lines=$(wc -l < gpurun_out/bench_n2.json)"; tail -5 gpurun_out/bench_n2.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_n2.json'))
    print('N=2 value %.4g e2e %.4g ms/step %.2f | per_rank %s | check %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['per_rank'], d['multi_gpu_check']))
    print('cfg3', d['config3_bq256']); print('two_stage', {k:v for k,v in (d['config4_two_stage'] or {}).items() if k in ('p50_ms','p95_ms','stage_ms_p50','recall','error')})
except Exception as e: print('parse failed', e)
PY
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_concurrency.py --sharded --clients 16 --rounds 10 --pages 32768 > gpurun_out/conc_n2.json 2> gpurun_out/conc_n2.err; echo "concurrency n2 rc=$?"; tail -c 1500 gpurun_out/conc_n2.json; tail -3 gpurun_out/conc_n2.err
