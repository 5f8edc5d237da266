#!/bin/bash
# round 2, visit A: fused-split GEMM correctness + speed
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_fused_gpu.py -x -q -m gpu > gpurun_out/a_fused_test.log 2>&1; echo "rc=$?" >> gpurun_out/a_fused_test.log
tail -n 30 gpurun_out/a_fused_test.log | cut -c1-300
for mode in 1 0; do
  TRN_GEMM_FUSED=$mode timeout 300 python scripts/exp/exp_fused.py > gpurun_out/a_fused_bench_$mode.log 2>&1; echo "rc=$?" >> gpurun_out/a_fused_bench_$mode.log
  cat gpurun_out/a_fused_bench_$mode.log | cut -c1-200
done
TRN_GEMM_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_pair -s 2 -c 1 -f -o gpurun_out/prof_fused_cfg3 python scripts/exp/exp_fused.py cfg3 > gpurun_out/a_prof.log 2>&1; echo "rc=$?" >> gpurun_out/a_prof.log
tail -n 5 gpurun_out/a_prof.log
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -x -q -m gpu -k "matmul or batched or gemm or config" > gpurun_out/a_parity.log 2>&1; echo "rc=$?" >> gpurun_out/a_parity.log
tail -n 15 gpurun_out/a_parity.log | cut -c1-300
