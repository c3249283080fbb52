set -x
mkdir -p gpurun_out/r2b
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2b/pytest_gpu.log 2>&1; tail -40 gpurun_out/r2b/pytest_gpu.log
timeout 600 python bench.py --micro > gpurun_out/r2b/micro.json 2> gpurun_out/r2b/micro.err; cat gpurun_out/r2b/micro.err | tail -30
