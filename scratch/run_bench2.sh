for P in bf16 tf32; do timeout 300 python bench.py --precision $P --steps 20 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/ktable_${P}m.txt > gpurun_out/bench_${P}m.json 2> gpurun_out/b.err; done
grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_tf32m.json gpurun_out/bench_bf16m.json
