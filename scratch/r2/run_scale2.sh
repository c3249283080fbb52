# final round-2 scaling evidence: run under gpurun --gpus 2 or --gpus 8
O=gpurun_out/scale2
mkdir -p $O
G=$(python -c "import torch; print(torch.cuda.device_count())")
B() { name=$1; n=$2; shift 2
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline "$@" > $O/$name.json 2> $O/$name.err
  python - <<PY
import json
try:
    d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1]); bv=d.get('baseline_variant') or {}
    print('$name', 'ms/step %.4f' % d['ms_per_step'], 'value %.0f' % d['value'], 'e2e', d['e2e']['value'], 'b64 ms', bv.get('ms_per_step'))
except Exception as e: print('$name FAILED', e)
PY
}
if [ "$G" -ge 8 ]; then
  B n8_sync 8
  B n8_nosync 8 --no-sync-bn
else
  timeout 500 python -m pytest tests/test_gpu_comm.py tests/test_gpu_ddp.py -m gpu -q --timeout 300 > $O/pytest_comm2.log 2>&1; tail -2 $O/pytest_comm2.log
  B n2_sync 2
fi
