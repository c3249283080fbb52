mkdir -p gpurun_out/r2o
timeout 150 python tests/multi_gpu/fit_n_gpus_check.py 2 > gpurun_out/r2o/fit2.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2o/fit2.log
timeout 500 python -m pytest tests/test_gpu_comm.py tests/test_gpu_ddp.py -m gpu -q --timeout 300 > gpurun_out/r2o/pytest_comm.log 2>&1; tail -4 gpurun_out/r2o/pytest_comm.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2o/bench_n2.json 2> gpurun_out/r2o/bench_n2.err; cut -c1-200 gpurun_out/r2o/bench_n2.json
