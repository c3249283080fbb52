mkdir -p gpurun_out/r2q
timeout 400 python -m pytest tests/test_gpu_ddp.py -m gpu -q --timeout 300 > gpurun_out/r2q/pytest_ddp.log 2>&1; tail -4 gpurun_out/r2q/pytest_ddp.log
for sp in 1 0; do
SN_BN_SPLIT_EXCHANGE=$sp timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2q/bench_n2_split$sp.json 2> gpurun_out/r2q/bench_n2_split$sp.err; tail -1 gpurun_out/r2q/bench_n2_split$sp.err; cut -c1-200 gpurun_out/r2q/bench_n2_split$sp.json
done
