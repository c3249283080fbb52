set -x
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/final/pytest_gpu.log
for P in bf16 tf32 fp32; do
  timeout 300 python bench.py --precision $P --steps 20 --warmup 3 --kernel-table gpurun_out/final/kernel_table_$P.json > gpurun_out/final/bench_${P}_n1.json 2> gpurun_out/final/bench_$P.err
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_reference.json 2>> gpurun_out/final/bench_ref.err
for P in bf16 tf32; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final/launches_$P.csv python bench.py --precision $P --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu_ll_$P.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stem_wgrad_kernel|stem_fprop_kernel|bn_pool_bwd_apply_kernel|bn_pool_bwd_reduce_pooled_kernel|bn_pool_fwd_kernel|conv_igemm_kernel|conv_wgrad_halo_kernel' -s 30 -c 26 -o gpurun_out/final/full_bf16 python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu_full.log 2>&1
tail -3 gpurun_out/final/pytest_gpu.log
cat gpurun_out/final/bench_bf16_n1.json
