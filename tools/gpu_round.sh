#!/bin/bash
# One gpurun call: parity tests, bench, ncu launch list of the bench command, full ncu capture of the hot kernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.csv
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python tools/profile_kernels.py --int8 --binary > gpurun_out/profile_plain.log 2>&1; echo "profile_plain rc=$?"; cat gpurun_out/profile_plain.log | tail -8
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'maxsim|topk|pack_rows|chunk_page|merge' -c 200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --pages 65536 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_umma -s 4 -c 2 -o gpurun_out/prof_umma_nm4 -f python tools/profile_kernels.py > gpurun_out/ncu_nm4.log 2>&1; echo "ncu nm4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_umma -s 10 -c 2 -o gpurun_out/prof_umma_nm1 -f python tools/profile_kernels.py > gpurun_out/ncu_nm1.log 2>&1; echo "ncu nm1 rc=$?"
ls -la gpurun_out
