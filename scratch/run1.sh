set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; echo "pytest exit $?" >> gpurun_out/t.log
timeout 300 python bench.py --precision tf32 --steps 20 --warmup 3 --kernel-table gpurun_out/ktable_tf32m.txt > gpurun_out/bench_tf32m.json 2> gpurun_out/b.err
timeout 300 python bench.py --precision bf16 --steps 20 --warmup 3 --kernel-table gpurun_out/ktable_bf16m.txt > gpurun_out/bench_bf16m.json 2>> gpurun_out/b.err
tail -3 gpurun_out/t.log; cat gpurun_out/bench_tf32m.json gpurun_out/bench_bf16m.json
