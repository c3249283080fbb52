mkdir -p gpurun_out/final
timeout 300 python bench.py --config cfg5 --steps 20 --warmup 3 --kernel-table gpurun_out/final/kernel_table_cfg5.json > gpurun_out/final/bench_cfg5.json 2> gpurun_out/final/bench_cfg5.err; tail -2 gpurun_out/final/bench_cfg5.err
timeout 300 python bench.py --config cfg5 --precision fp32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/final/bench_cfg5_fp32.json 2> gpurun_out/final/bench_cfg5_fp32.err
cat gpurun_out/final/bench_cfg5.json | cut -c1-1500
