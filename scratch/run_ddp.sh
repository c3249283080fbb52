timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu/ddp_check.py > gpurun_out/ddp.log 2>&1
echo "exit $?"
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/ddp.log | grep -v "^$" | tail -25
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --config cfg5 --steps 20 --warmup 3 > gpurun_out/bench_cfg5_n2.json 2> gpurun_out/bench_cfg5_n2.err; tail -2 gpurun_out/bench_cfg5_n2.err; cut -c1-400 gpurun_out/bench_cfg5_n2.json
