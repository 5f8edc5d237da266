#!/bin/bash
# final evidence pass of round 2 (one GPU): the changed row / reduction kernels under ncu --set full, the bench launch list
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "ring_kernel or row_blocks or pipelined" 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r02_rows python scripts/profile_kernels.py softmax map slice > gpurun_out/prof_r02_rows.log 2>&1; echo "rc=$?" >> gpurun_out/prof_r02_rows.log; tail -n 2 gpurun_out/prof_r02_rows.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/prof_r02_rows.ncu-rep gpurun_out/r02_bench_launches.csv
