N=$1
export SN_BENCH_WATCHDOG=240
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --sync-bn > gpurun_out/bench_bf16_syncbn_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit $?"
tail -2 gpurun_out/bench_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_bf16_syncbn_n$N.json').read().strip().splitlines()[-1]); print('sync_bn N=$N value', round(d['value']), 'ms', d['ms_per_step'], 'e2e', round(d['e2e']['value']))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_bf16_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_bf16_n$N.json').read().strip().splitlines()[-1]); print('plain  N=$N value', round(d['value']), 'ms', d['ms_per_step'], 'e2e', round(d['e2e']['value']))"
