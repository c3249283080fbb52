timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 300 python bench.py --precision tf32x3 --steps 10 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/ktable_tf32x3.json > gpurun_out/bench_tf32x3.json 2> gpurun_out/b.err; tail -3 gpurun_out/b.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_tf32x3.json').read().strip().splitlines()[-1]); print('tf32x3 value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches']//d['steps'])"
