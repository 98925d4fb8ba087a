#!/bin/bash
set -u
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
echo "== w4 on"; timeout 600 python tools/profile_kernels.py --int8 --pages 65536 2>&1 | tail -4 | cut -c1-90
echo "== w4 off"; B200MS_EPI_W4=0 timeout 600 python tools/profile_kernels.py --int8 --pages 65536 2>&1 | tail -4 | cut -c1-90
