mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --precision bf16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/b.err; tail -3 gpurun_out/b.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_bf16.json').read().strip().splitlines()[-1]); print('bf16 value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches']//d['steps'], d['roofline']['frac'])"
