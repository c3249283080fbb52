mkdir -p gpurun_out/r2r
timeout 400 python bench.py --steps 20 --warmup 5 --kernel-table gpurun_out/r2r/kt_bf16.json > gpurun_out/r2r/bench_n1.json 2> gpurun_out/r2r/bench_n1.err; tail -3 gpurun_out/r2r/bench_n1.err; cat gpurun_out/r2r/bench_n1.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2r/bench_ref.json 2> gpurun_out/r2r/bench_ref.err; cut -c1-700 gpurun_out/r2r/bench_ref.json
