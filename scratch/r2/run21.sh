mkdir -p gpurun_out/r2u
for r in 1 0 1 0; do SN_IGEMM_ROWS=$r timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rows=$r ms/step %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"; done
timeout 300 python bench.py --config cfg4 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200
SN_IGEMM_ROWS=0 timeout 300 python bench.py --config cfg4 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-200
