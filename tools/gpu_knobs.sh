#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_store.py -q -x 2>&1 | tail -2
for ur in 1024 4096 16384 65536; do echo "== unit_rows $ur"; B200MS_UNIT_ROWS=$ur timeout 600 python tools/profile_kernels.py --int8 --pages 65536 2>&1 | tail -4 | cut -c1-90; done
