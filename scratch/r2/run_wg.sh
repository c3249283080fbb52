# A/B of the TMA reduce-add wgrad epilogues (SN_WGRAD_REDG=1: the old red.global / atomicAdd epilogue) + deeper stem patch ring
O=gpurun_out/wg
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tf32.py tests/test_gpu_fullsize.py tests/test_gpu_ops.py -m gpu -q -x --timeout 600 -k "wgrad or dense or Dense or block or trajectory or conv" > $O/pytest_wg.log 2>&1; echo "pytest exit $?" >> $O/pytest_wg.log; tail -4 $O/pytest_wg.log
for red in 0 1; do
  SN_WGRAD_REDG=$red timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernel-table $O/kt_bf16_redg$red.json > $O/bench_bf16_redg$red.json 2> $O/bench_redg$red.err; cut -c1-120 $O/bench_bf16_redg$red.json
  SN_WGRAD_REDG=$red timeout 300 python bench.py --config cfg5 --steps 20 --warmup 5 --no-cpu-baseline --kernel-table $O/kt_cfg5_redg$red.json > $O/bench_cfg5_redg$red.json 2> $O/bench_cfg5_redg$red.err; cut -c1-120 $O/bench_cfg5_redg$red.json
done
python - <<'P'
import json
for red in (0,1):
    for cfg in ('bf16','cfg5'):
        try:
            kt=json.load(open('gpurun_out/wg/kt_%s_redg%d.json'%(cfg,red)))
            print(cfg,'redg',red,'step',kt['step_ms_graph'])
            for r in kt['kernels']:
                if 'wgrad' in r['name']: print('   ',r['name'],r['args'][:6],round(r['ms']*1e3,1),'us')
        except Exception as e: print(cfg,red,'ERR',e)
P
