mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/final/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/final/smoke.log
for P in bf16 tf32; do
  timeout 300 python bench.py --precision $P --steps 20 --warmup 3 --kernel-table gpurun_out/final/kernel_table_$P.json > gpurun_out/final/bench_${P}_n1.json 2> gpurun_out/final/bench_$P.err
done
timeout 300 python bench.py --config cfg4 --steps 20 --warmup 3 --kernel-table gpurun_out/final/kernel_table_cfg4.json > gpurun_out/final/bench_cfg4.json 2> gpurun_out/final/bench_cfg4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final/launches_bf16.csv python bench.py --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu_ll_bf16.log 2>&1
tail -2 gpurun_out/final/pytest_gpu.log; tail -2 gpurun_out/final/smoke.log
for f in bf16_n1 tf32_n1 cfg4; do python -c "
import json,sys; d=json.loads(open('gpurun_out/final/bench_%s.json'%sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches']//d['steps'], 'roof', d.get('roofline',{}).get('frac'), d['clocks'])" $f; done
