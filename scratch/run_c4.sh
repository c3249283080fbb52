mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_ops.py -q -x 2>&1 | tail -8
timeout 300 python bench.py --config cfg4 --steps 20 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/ktable_cfg4.json > gpurun_out/bench_cfg4.json 2> gpurun_out/b4.err; tail -3 gpurun_out/b4.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg4.json').read().strip().splitlines()[-1]); print('cfg4 value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches']//d['steps'])
k=json.load(open('gpurun_out/ktable_cfg4.json'))
agg={}
for r in k['kernels']: agg[r['name']]=agg.get(r['name'],0)+r['ms']*r['launches']
tot=sum(agg.values())
for n,v in sorted(agg.items(), key=lambda kv:-kv[1])[:12]: print('  %-28s %5.1f%%'%(n,100*v/tot))"
