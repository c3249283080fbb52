O=gpurun_out/lstm
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_recurrent.py -m gpu -q --timeout 300 > $O/pytest_lstm.log 2>&1; echo "pytest exit $?" >> $O/pytest_lstm.log; tail -8 $O/pytest_lstm.log
timeout 300 python bench.py --config cfg5 --steps 20 --warmup 5 --no-cpu-baseline --kernel-table $O/kt_cfg5.json > $O/bench_cfg5.json 2> $O/bench_cfg5.err; cut -c1-120 $O/bench_cfg5.json; tail -2 $O/bench_cfg5.err
python - <<'P'
import json
kt=json.load(open('gpurun_out/lstm/kt_cfg5.json'))
print('step',kt['step_ms_graph'])
for r in kt['kernels'][:4]: print('   ',r['name'],r['args'][:5],round(r['ms']*1e3,1),'us')
P
