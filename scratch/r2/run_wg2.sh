# bf16 stem wgrad (dY as the bf16 map) + three builder groups: parity tests, then A/B of the step
O=gpurun_out/wg2
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tf32.py tests/test_gpu_fullsize.py tests/test_gpu_train.py -m gpu -q -x --timeout 600 > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -5 $O/pytest.log
for v in 1 0; do
  SN_STEM_WGRAD_BF16=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernel-table $O/kt_bf16_stem$v.json > $O/bench_bf16_stem$v.json 2> $O/bench_stem$v.err; cut -c1-120 $O/bench_bf16_stem$v.json
done
python - <<'P'
import json
for v in (1,0):
    try:
        kt=json.load(open('gpurun_out/wg2/kt_bf16_stem%d.json'%v))
        print('stem16',v,'step',kt['step_ms_graph'])
        for r in kt['kernels'][:12]:
            print('   ',r['name'],r['args'][:5],round(r['ms']*1e3,1),'us')
    except Exception as e: print(v,'ERR',e)
P
