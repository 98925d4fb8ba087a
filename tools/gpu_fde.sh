#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_fde.py -q -x 2>&1 | tail -25
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fde.py 2>&1 | tail -5
echo "== sanitizer (memcheck) on smoke"; timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py smoke 2>&1 | tail -8
echo "== sanitizer (racecheck) on smoke"; timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python __graft_entry__.py smoke 2>&1 | tail -8
