#!/bin/bash
# A/B of the SS (A from smem) and TS (A from TMEM) forms of maxsim_umma: parity tests + kernel times for both.
set -u
mkdir -p gpurun_out
echo "== TS parity"; B200MS_A_IN_TMEM=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_store.py -q -x 2>&1 | tail -12
echo "== SS parity"; B200MS_A_IN_TMEM=0 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_store.py -q -x 2>&1 | tail -4
echo "== TS times"; B200MS_A_IN_TMEM=1 timeout 600 python tools/profile_kernels.py --int8 --binary --pages 65536 2>&1 | tail -8
echo "== SS times"; B200MS_A_IN_TMEM=0 timeout 600 python tools/profile_kernels.py --int8 --pages 65536 2>&1 | tail -8
