#!/bin/bash
# One GPU-box visit.  Usage: gpu_round.sh [quick|full|prof]
mkdir -p gpurun_out
mode=${1:-full}
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_vector_ext_gpu.py tests/test_vector_api_gpu.py -q -m gpu -k "softmax or layer_norm or vector_api" > gpurun_out/t_softmax.log 2>&1; echo "rc=$?" >> gpurun_out/t_softmax.log
timeout 300 python scripts/sweep_rows_ext.py > gpurun_out/sweep_ext_new.log 2>&1
if [ "$mode" = quick ]; then
timeout 600 python scripts/exp/exp_long_rows.py > gpurun_out/exp_long_rows.log 2>&1
fi
if [ "$mode" = prof ] || [ "$mode" = final ]; then
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_rows2 python scripts/profile_kernels.py rows2 > gpurun_out/prof_rows2.log 2>&1
fi
if [ "$mode" = final ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
fi
if [ "$mode" = full ] || [ "$mode" = final ]; then
timeout 1200 python -m pytest tests -q -m gpu --durations=15 > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
fi
for f in t_softmax t_all smoke prof_rows2; do [ -f gpurun_out/$f.log ] && tail -n 25 gpurun_out/$f.log | cut -c1-200; done; cat gpurun_out/sweep_ext_new.log; [ "$mode" = quick ] && cat gpurun_out/exp_long_rows.log; true
