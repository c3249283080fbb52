mkdir -p gpurun_out/r2g
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_ddp.py -m gpu -q -x --timeout 300 > gpurun_out/r2g/pytest_comm.log 2>&1; tail -5 gpurun_out/r2g/pytest_comm.log
for fuse in 1 0; do
SN_FUSE_SGD=$fuse timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2g/bench_n2_fuse$fuse.json 2> gpurun_out/r2g/bench_n2_fuse$fuse.err; tail -2 gpurun_out/r2g/bench_n2_fuse$fuse.err; cat gpurun_out/r2g/bench_n2_fuse$fuse.json | cut -c1-200
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-sync-bn > gpurun_out/r2g/bench_n2_nosync.json 2> gpurun_out/r2g/bench_n2_nosync.err; cat gpurun_out/r2g/bench_n2_nosync.json | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --per-gpu-batch 64 > gpurun_out/r2g/bench_n2_b64.json 2> gpurun_out/r2g/bench_n2_b64.err; cat gpurun_out/r2g/bench_n2_b64.json | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --per-gpu-batch 64 > gpurun_out/r2g/bench_n1_b64.json 2> gpurun_out/r2g/bench_n1_b64.err; cat gpurun_out/r2g/bench_n1_b64.json | cut -c1-200
