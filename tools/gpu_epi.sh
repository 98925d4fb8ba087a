#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -2
timeout 600 python tools/profile_kernels.py --int8 --pages 65536 2>&1 | tail -4
