N=$1
export SN_BENCH_WATCHDOG=240
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_bf16_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit $?"
tail -2 gpurun_out/bench_n$N.err
cat gpurun_out/bench_bf16_n$N.json
