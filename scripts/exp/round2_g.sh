#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 > gpurun_out/g_all.log 2>&1; echo "rc=$?" >> gpurun_out/g_all.log
tail -n 22 gpurun_out/g_all.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"
python scripts/exp/show_bench.py gpurun_out/g_bench.json
