mkdir -p gpurun_out/r2h
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x --timeout 600 > gpurun_out/r2h/fullsize.log 2>&1; tail -25 gpurun_out/r2h/fullsize.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernel-table gpurun_out/r2h/kt_bf16maps.json > gpurun_out/r2h/bench_maps1.json 2> gpurun_out/r2h/bench_maps1.err; tail -3 gpurun_out/r2h/bench_maps1.err; cut -c1-1500 gpurun_out/r2h/bench_maps1.json
SN_BF16_MAPS=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h/bench_maps0.json 2> gpurun_out/r2h/bench_maps0.err; cut -c1-200 gpurun_out/r2h/bench_maps0.json
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_fullsize.py > gpurun_out/r2h/pytest_gpu.log 2>&1; tail -15 gpurun_out/r2h/pytest_gpu.log
