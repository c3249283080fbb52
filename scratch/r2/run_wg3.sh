# stem wgrad with lane-parallel patch geometry; LSTM recurrence on the tensor pipe (tests, A/B of config 5)
O=gpurun_out/wg3
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_recurrent.py -m gpu -q --timeout 300 > $O/pytest_lstm.log 2>&1; echo "pytest exit $?" >> $O/pytest_lstm.log; tail -15 $O/pytest_lstm.log
timeout 600 python -m pytest tests/test_gpu_tf32.py -m gpu -q -x --timeout 300 -k "stem or conv2d_tf32 or trajectory" > $O/pytest_stem.log 2>&1; echo "pytest exit $?" >> $O/pytest_stem.log; tail -3 $O/pytest_stem.log
for v in 1 0; do
  SN_LSTM_TC=$v timeout 300 python bench.py --config cfg5 --steps 20 --warmup 5 --no-cpu-baseline --kernel-table $O/kt_cfg5_tc$v.json > $O/bench_cfg5_tc$v.json 2> $O/bench_cfg5_tc$v.err; cut -c1-120 $O/bench_cfg5_tc$v.json; tail -2 $O/bench_cfg5_tc$v.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernel-table $O/kt_bf16.json > $O/bench_bf16.json 2> $O/bench_bf16.err; cut -c1-120 $O/bench_bf16.json
python - <<'P'
import json
for f in ('kt_cfg5_tc1','kt_cfg5_tc0','kt_bf16'):
    try:
        kt=json.load(open('gpurun_out/wg3/%s.json'%f))
        print(f,'step',kt['step_ms_graph'])
        for r in kt['kernels'][:6]:
            print('   ',r['name'],r['args'][:5],round(r['ms']*1e3,1),'us')
    except Exception as e: print(f,'ERR',e)
P
