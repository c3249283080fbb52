mkdir -p gpurun_out/r2m
for cm in 0 20971520; do
SN_BNPOOL_COOP_MAX=$cm timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m/bench_n2_coop$cm.json 2> gpurun_out/r2m/bench_n2_coop$cm.err; tail -2 gpurun_out/r2m/bench_n2_coop$cm.err; cat gpurun_out/r2m/bench_n2_coop$cm.json | cut -c1-200
done
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_ddp.py -m gpu -q -x --timeout 300 > gpurun_out/r2m/pytest_comm.log 2>&1; tail -5 gpurun_out/r2m/pytest_comm.log
