mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --precision bf16 --steps 20 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/ktable_bf16.json > gpurun_out/bench_bf16.json 2> gpurun_out/b.err; tail -3 gpurun_out/b.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_bf16.json').read().strip().splitlines()[-1]); print('bf16 value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches']//d['steps'], d['roofline']['frac'], d['roofline']['kernel_ms'])
k=json.load(open('gpurun_out/ktable_bf16.json'))
for r in sorted(k['kernels'], key=lambda r:-r['ms']*r['launches']):
    if 'finalize' in r['name'] or 'bn_pool_bwd' in r['name']: print('  ', r['name'], r['args'][:4], round(r['ms']*1e3,1), 'us')"
