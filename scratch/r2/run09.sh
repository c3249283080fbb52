mkdir -p gpurun_out/r2i
timeout 300 python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2i/plain.json 2> gpurun_out/r2i/plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stem_wgrad_kernel|stem_fprop_kernel|bn_pool_bwd_apply_kernel|bn_pool_bwd_reduce_pooled_kernel|bn_pool_fwd_kernel|conv_igemm_kernel|conv_wgrad_halo_kernel' -s 30 -c 24 -o gpurun_out/r2i/full_bf16 python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2i/ncu_full.log 2>&1
ls -la gpurun_out/r2i; tail -3 gpurun_out/r2i/ncu_full.log
