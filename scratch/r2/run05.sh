mkdir -p gpurun_out/r2e
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_ddp.py -m gpu -q -x --timeout 300 > gpurun_out/r2e/pytest_comm.log 2>&1; tail -5 gpurun_out/r2e/pytest_comm.log
for ov in 1 0; do
SN_COMM_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e/bench_n2_ov$ov.json 2> gpurun_out/r2e/bench_n2_ov$ov.err; tail -2 gpurun_out/r2e/bench_n2_ov$ov.err; cat gpurun_out/r2e/bench_n2_ov$ov.json | cut -c1-200
SN_COMM_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-sync-bn > gpurun_out/r2e/bench_n2_nosync_ov$ov.json 2> gpurun_out/r2e/bench_n2_nosync_ov$ov.err; cat gpurun_out/r2e/bench_n2_nosync_ov$ov.json | cut -c1-200
done
