import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', round(d['e2e']['value'],1))
for s in d['secondary']:
    print(s['key'].ljust(16), round(s['ms']*1e3,1), 'us  one_gpu', s.get('ms_one_gpu') and round(s['ms_one_gpu']*1e3,1), 'per rank', [round(v*1e3,1) for v in s.get('ms_per_rank',[])])
print('parity', d['parity']); print('strong', d.get('strong_scaling'))
