set -x
mkdir -p gpurun_out/r2a
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r2a/smoke.log 2>&1; tail -3 gpurun_out/r2a/smoke.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x --timeout 600 > gpurun_out/r2a/fullsize.log 2>&1; tail -30 gpurun_out/r2a/fullsize.log
timeout 300 python bench.py --steps 20 --warmup 3 --kernel-table gpurun_out/r2a/ktable_bf16.json > gpurun_out/r2a/bench_bf16.json 2> gpurun_out/r2a/bench.err; cat gpurun_out/r2a/bench_bf16.json
