#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_fused_gpu.py -x -q -m gpu > gpurun_out/c_fused_test.log 2>&1; echo "rc=$?" >> gpurun_out/c_fused_test.log
tail -n 12 gpurun_out/c_fused_test.log | cut -c1-300
TRN_GEMM_FUSED=1 timeout 200 python scripts/exp/exp_fused.py cfg3 2>&1 | tail -3
TRN_GEMM_FUSED=1 TRN_GEMM_ASTAT=0 timeout 200 python scripts/exp/exp_fused.py cfg3 2>&1 | tail -3
TRN_GEMM_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_astat -s 2 -c 1 -f -o gpurun_out/prof_astat_cfg3 python scripts/exp/exp_fused.py cfg3 > gpurun_out/c_prof.log 2>&1; echo "rc=$?" >> gpurun_out/c_prof.log
tail -n 3 gpurun_out/c_prof.log
