mkdir -p gpurun_out/r2l
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2l/pytest_gpu.log 2>&1; tail -12 gpurun_out/r2l/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernel-table gpurun_out/r2l/kt.json > gpurun_out/r2l/bench.json 2> gpurun_out/r2l/bench.err; tail -3 gpurun_out/r2l/bench.err; cut -c1-200 gpurun_out/r2l/bench.json
SN_BNPOOL_COOP_MAX=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2l/bench_nocoop.json 2> gpurun_out/r2l/bench_nocoop.err; cut -c1-200 gpurun_out/r2l/bench_nocoop.json
SN_BNPOOL_COOP_MAX=100000000 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2l/bench_allcoop.json 2> gpurun_out/r2l/bench_allcoop.err; cut -c1-200 gpurun_out/r2l/bench_allcoop.json
