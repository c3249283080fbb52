for K in 20 100; do
timeout 300 python bench.py --steps $K --warmup 3 --no-cpu-baseline > gpurun_out/bench_k$K.json 2> gpurun_out/b.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_k$K.json').read().strip().splitlines()[-1]); print('K=$K value ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4))"
done
