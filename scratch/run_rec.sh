mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_recurrent.py -q 2>&1 | tail -5
timeout 300 python bench.py --config cfg5 --precision fp32 --steps 10 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/ktable_cfg5.json > gpurun_out/bench_cfg5.json 2> gpurun_out/b5.err; tail -3 gpurun_out/b5.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg5.json').read().strip().splitlines()[-1]); print('cfg5 value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches']//d['steps'])
k=json.load(open('gpurun_out/ktable_cfg5.json'))
for r in sorted(k['kernels'], key=lambda r:-r['ms']*r['launches'])[:6]: print(r['name'], r['args'], round(r['ms']*1e3,1), 'us x', r['launches'])"
