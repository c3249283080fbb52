mkdir -p gpurun_out
for p in tf32 bf16; do
timeout 300 python bench.py --config cfg5 --precision $p --steps 20 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/ktable_cfg5_$p.json > gpurun_out/bench_cfg5_$p.json 2> gpurun_out/b5.err; tail -3 gpurun_out/b5.err
python -c "
import json,sys; p=sys.argv[1]; d=json.loads(open('gpurun_out/bench_cfg5_%s.json'%p).read().strip().splitlines()[-1]); print(p, 'cfg5 value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches']//d['steps'], 'loss', d['e2e'].get('final_loss'))
k=json.load(open('gpurun_out/ktable_cfg5_%s.json'%p))
for r in sorted(k['kernels'], key=lambda r:-r['ms']*r['launches'])[:8]: print('  ', r['name'], r['args'], round(r['ms']*1e3,1), 'us x', r['launches'])" $p
done
