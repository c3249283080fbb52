O=gpurun_out/bnp
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 600 > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernel-table $O/kt_bf16.json > $O/bench_bf16.json 2> $O/bench_bf16.err; cut -c1-120 $O/bench_bf16.json
python - <<'P'
import json
kt=json.load(open('gpurun_out/bnp/kt_bf16.json'))
print('step',kt['step_ms_graph'])
for r in kt['kernels'][:14]: print('   ',r['name'],r['args'][:5],round(r['ms']*1e3,1),'us')
P
