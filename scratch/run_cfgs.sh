for C in cfg1 cfg2 cfg4; do
  timeout 300 python bench.py --config $C --steps 20 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/ktable_$C.json > gpurun_out/bench_$C.json 2> gpurun_out/bench_$C.err; echo "$C exit $?"
  python - "$C" <<'PY'
import json,sys
c=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/bench_%s.json'%c).read().strip().splitlines()[-1])
    print(c, 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches/step', d['gpu_launches']//d['steps'], d['roofline']['kernel'], round(d['roofline']['frac'],3))
except Exception as e:
    print(c, 'failed', e); print(open('gpurun_out/bench_%s.err'%c).read()[-1500:])
PY
done
