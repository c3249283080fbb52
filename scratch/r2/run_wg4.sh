O=gpurun_out/wg4
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_recurrent.py -m gpu -q --timeout 300 > $O/pytest_lstm.log 2>&1; echo "pytest exit $?" >> $O/pytest_lstm.log; tail -4 $O/pytest_lstm.log
timeout 600 python -m pytest tests/test_gpu_tf32.py -m gpu -q -x --timeout 300 -k "stem or conv2d_tf32 or trajectory" > $O/pytest_stem.log 2>&1; echo "pytest exit $?" >> $O/pytest_stem.log; tail -3 $O/pytest_stem.log
timeout 300 python bench.py --config cfg5 --steps 20 --warmup 5 --no-cpu-baseline --kernel-table $O/kt_cfg5.json > $O/bench_cfg5.json 2> $O/bench_cfg5.err; cut -c1-120 $O/bench_cfg5.json; tail -2 $O/bench_cfg5.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernel-table $O/kt_bf16.json > $O/bench_bf16.json 2> $O/bench_bf16.err; cut -c1-120 $O/bench_bf16.json
python - <<'P'
import json
for f in ('kt_cfg5','kt_bf16'):
    try:
        kt=json.load(open('gpurun_out/wg4/%s.json'%f))
        print(f,'step',kt['step_ms_graph'])
        for r in kt['kernels'][:5]:
            print('   ',r['name'],r['args'][:5],round(r['ms']*1e3,1),'us')
    except Exception as e: print(f,'ERR',e)
P
