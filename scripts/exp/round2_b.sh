#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_fused_gpu.py -x -q -m gpu > gpurun_out/b_fused_test.log 2>&1; echo "rc=$?" >> gpurun_out/b_fused_test.log
tail -n 12 gpurun_out/b_fused_test.log | cut -c1-300
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 100 > gpurun_out/b_clocks.csv &
SMI=$!
TRN_GEMM_FUSED=1 timeout 200 python scripts/exp/exp_fused.py cfg3 2>&1 | tail -3
kill $SMI
sort gpurun_out/b_clocks.csv | uniq -c | sort -rn | head -5
TRN_GEMM_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_pair -s 2 -c 1 -f -o gpurun_out/prof_fused_cfg3 python scripts/exp/exp_fused.py cfg3 > gpurun_out/b_prof.log 2>&1; echo "rc=$?" >> gpurun_out/b_prof.log
tail -n 3 gpurun_out/b_prof.log
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -x -q -m gpu -k "matmul or batched or gemm or config" > gpurun_out/b_parity.log 2>&1; echo "rc=$?" >> gpurun_out/b_parity.log
tail -n 8 gpurun_out/b_parity.log | cut -c1-300
