#!/bin/bash
set -u
echo "== parity, tensor 1-bit path (default)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_store.py tests/test_gpu_fde.py tests/test_gpu_shardfile_reranker.py -q -x 2>&1 | tail -8
echo "== parity, popc path"; B200MS_B1_TENSOR=0 timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "binary or topk" 2>&1 | tail -3
echo "== times tensor"; timeout 600 python tools/profile_kernels.py --binary --pages 65536 2>&1 | tail -3
echo "== times popc"; B200MS_B1_TENSOR=0 timeout 600 python tools/profile_kernels.py --binary --pages 65536 2>&1 | tail -3
echo "== two stage"; timeout 600 python tools/bench_two_stage.py --pages 65536
