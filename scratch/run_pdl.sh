mkdir -p gpurun_out
SN_PDL=0 timeout 200 python bench.py --precision bf16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl0.json 2> gpurun_out/pdl0.err; echo "pdl0 exit $?"
SN_PDL=1 timeout 200 python bench.py --precision bf16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl1.json 2> gpurun_out/pdl1.err; echo "pdl1 exit $?"; tail -3 gpurun_out/pdl1.err
for m in 0 1; do python -c "
import json,sys; d=json.loads(open('gpurun_out/bench_pdl%s.json'%sys.argv[1]).read().strip().splitlines()[-1]); print('SN_PDL=%s'%sys.argv[1], round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'loss', d['e2e']['final_loss'])" $m; done
SN_PDL=1 timeout 400 python -m pytest tests/test_gpu_train.py tests/test_gpu_tf32.py -q -x 2>&1 | tail -4
