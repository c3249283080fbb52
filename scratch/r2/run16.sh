mkdir -p gpurun_out/r2p
timeout 200 python -m pytest tests/test_gpu_comm.py -m gpu -q --timeout 150 > gpurun_out/r2p/pytest_comm8.log 2>&1; tail -3 gpurun_out/r2p/pytest_comm8.log
B() { # name nproc extra...
  name=$1; n=$2; shift 2
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2p/$name.json 2> gpurun_out/r2p/$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2p/$name.json').read().strip().splitlines()[-1]); print('$name', 'ms/step %.4f' % d['ms_per_step'], 'value %.0f' % d['value'], 'e2e ms %.4f' % d['e2e']['ms_per_step'])
except Exception as e: print('$name FAILED', e)
PY
}
B n8_sync 8
B n8_nosync 8 --no-sync-bn
B n8_b64 8 --per-gpu-batch 64
B n4_sync 4
SN_COMM=nccl B n8_nccl_sync 8
timeout 100 python tests/multi_gpu/fit_n_gpus_check.py 8 > gpurun_out/r2p/fit8.log 2>&1; tail -3 gpurun_out/r2p/fit8.log
