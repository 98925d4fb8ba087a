#!/bin/bash
set -u
B200MS_B1_TENSOR=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "binary or topk" 2>&1 | tail -2
echo "== tensor"; B200MS_B1_TENSOR=1 timeout 600 python tools/profile_kernels.py --binary --pages 65536 2>&1 | tail -2
