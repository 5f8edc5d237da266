#!/bin/bash
# ncu evidence of round 2 (single GPU): --set full captures of the new / changed kernels, and the launch list of bench.py
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r02_gemm python scripts/profile_kernels.py gemm batched fused > gpurun_out/prof_r02_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/prof_r02_gemm.log; tail -n 2 gpurun_out/prof_r02_gemm.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r02_rows python scripts/profile_kernels.py softmax map slice > gpurun_out/prof_r02_rows.log 2>&1; echo "rc=$?" >> gpurun_out/prof_r02_rows.log; tail -n 2 gpurun_out/prof_r02_rows.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_bench_launches.csv
