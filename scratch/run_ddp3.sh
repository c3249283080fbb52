mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu/ddp_check.py > gpurun_out/ddp.log 2>&1
echo "exit $?"; grep -c "DDP_CHECK_OK" gpurun_out/ddp.log; grep "max rel err" gpurun_out/ddp.log | cut -c1-160
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_cfg3_n2.json 2> gpurun_out/bench_n2.err; echo "bench exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_n2.json').read().strip().splitlines()[-1]); print('cfg3 N=2', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'loss', d['e2e'].get('final_loss'))"
