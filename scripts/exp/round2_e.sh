#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_peer_loopback_gpu.py -x -q -m gpu > gpurun_out/e_peer.log 2>&1; echo "rc=$?" >> gpurun_out/e_peer.log
tail -n 25 gpurun_out/e_peer.log | cut -c1-250
( time timeout 900 python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err ) 2>&1 | grep real; echo "rc=$?"
tail -n 5 gpurun_out/e_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/e_bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('metric','value','ms_per_step','gpu_launches')})
    print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    print('roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['prepass_ms'])
    print('cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
    for s in d['secondary']:
        print(s['key'], round(s['value'],1), round(s['ms'],4), round(s['roofline_frac'],3), 'cpu', s.get('cpu_baseline',{}).get('value'), 'e2e', s.get('e2e',{}).get('value'))
    print('config1', json.dumps(d.get('config1'))[:900])
    print('parity', d['parity'])
    print('clocks', d['clocks'])
except Exception as e:
    print('parse failed', e)
PY
