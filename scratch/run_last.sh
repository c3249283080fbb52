mkdir -p gpurun_out/final
timeout 200 python bench.py --precision bf16 --steps 20 --warmup 3 --kernel-table gpurun_out/final/kernel_table_bf16.json > gpurun_out/final/bench_bf16_n1.json 2> gpurun_out/final/bench_bf16.err; echo "exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/final/bench_bf16_n1.json').read().strip().splitlines()[-1]); print('bf16', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'roof', d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['share_of_step'], 'cpu', d['cpu_baseline']['value'], d['clocks'])"
