mkdir -p gpurun_out/r2n
timeout 900 python -m pytest tests/test_gpu_tf32.py tests/test_gpu_ops.py tests/test_gpu_train.py -m gpu -q --timeout 600 -x > gpurun_out/r2n/pytest.log 2>&1; tail -12 gpurun_out/r2n/pytest.log
timeout 300 python bench.py --config cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n/bench_cfg2.json 2> gpurun_out/r2n/bench_cfg2.err; tail -2 gpurun_out/r2n/bench_cfg2.err; cut -c1-200 gpurun_out/r2n/bench_cfg2.json
timeout 600 python bench.py --micro > gpurun_out/r2n/micro.json 2> gpurun_out/r2n/micro.err; grep -i "dense\|adaptive" gpurun_out/r2n/micro.err
