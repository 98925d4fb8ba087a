#!/bin/bash
set -u
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "== split4 on"; timeout 600 python tools/profile_kernels.py --int8 --pages 65536 2>&1 | tail -4
echo "== split4 off"; B200MS_SPLIT4=0 timeout 600 python tools/profile_kernels.py --int8 --pages 65536 2>&1 | tail -4
