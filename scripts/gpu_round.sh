#!/bin/bash
# One GPU-box visit.  Usage: gpu_round.sh [quick|full]
mkdir -p gpurun_out
mode=${1:-full}
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_vector_ext_gpu.py -q -m gpu -k "softmax or layer_norm" > gpurun_out/t_softmax.log 2>&1; echo "rc=$?" >> gpurun_out/t_softmax.log
timeout 300 python scripts/sweep_rows_ext.py > gpurun_out/sweep_ext_new.log 2>&1
timeout 600 python scripts/exp/exp_long_rows.py > gpurun_out/exp_long_rows.log 2>&1
if [ "$mode" = full ]; then
timeout 1200 python -m pytest tests -q -m gpu --durations=15 > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
fi
for f in t_softmax t_all smoke; do tail -n 4 gpurun_out/$f.log; done; cat gpurun_out/sweep_ext_new.log gpurun_out/exp_long_rows.log
