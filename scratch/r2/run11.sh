mkdir -p gpurun_out/r2k
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2k/pytest_gpu.log 2>&1; tail -8 gpurun_out/r2k/pytest_gpu.log
