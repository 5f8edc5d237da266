#!/bin/bash
# 2-GPU visit: sharded-path parity under torchrun, then the strong-scaled bench
mkdir -p gpurun_out
NG=${1:-2}
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_peer_loopback_gpu.py -x -q -m gpu > gpurun_out/f_dist.log 2>&1; echo "rc=$?" >> gpurun_out/f_dist.log
tail -n 25 gpurun_out/f_dist.log | cut -c1-300
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/f_bench_$NG.json 2> gpurun_out/f_bench_$NG.err ) 2>&1 | grep real; echo "rc=$?"
tail -n 12 gpurun_out/f_bench_$NG.err | cut -c1-300
python - $NG <<'PY'
import json, sys
ng = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/f_bench_{ng}.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('metric','value','ms_per_step','gpu_launches','n_gpus')})
    print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    print('roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['prepass_ms'], d['roofline']['peak'])
    for s in d['secondary']:
        print(s['key'], round(s['value'],1), round(s['ms'],4), round(s['roofline_frac'],3), 'one_gpu_ms', s.get('ms_one_gpu'))
    print('parity', d['parity'])
    print('strong', d.get('strong_scaling'))
    print('clocks', d['clocks'])
except Exception as e:
    print('parse failed', e)
PY
