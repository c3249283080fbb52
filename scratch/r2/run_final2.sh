O=gpurun_out/final
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 --kernel-table $O/kernel_table_bf16.json > $O/bench_bf16_n1.json 2> $O/bench_bf16.err; cut -c1-250 $O/bench_bf16_n1.json
timeout 300 python bench.py --config cfg4 --steps 20 --warmup 5 --kernel-table $O/kernel_table_cfg4.json > $O/bench_cfg4.json 2> $O/bench_cfg4.err; cut -c1-200 $O/bench_cfg4.json
timeout 300 python bench.py --micro > $O/micro.json 2> $O/micro.err; tail -5 $O/micro.err
