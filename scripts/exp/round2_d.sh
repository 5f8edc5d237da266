#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_fused_gpu.py tests/test_parity_gpu.py tests/test_full_size_gpu.py tests/test_batch_gpu.py tests/test_attention_gpu.py -x -q -m gpu > gpurun_out/d_parity.log 2>&1; echo "rc=$?" >> gpurun_out/d_parity.log
tail -n 8 gpurun_out/d_parity.log | cut -c1-300
timeout 200 python scripts/exp/exp_fused.py 2>&1 | tail -12
