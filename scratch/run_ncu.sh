set -x
P=${1:-tf32}
# launch list of the whole bench (cold-cache per-launch times; shares must agree with the event table)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01c_$P.csv python bench.py --precision $P --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1
# full captures of the top kernels (one launch each, skipping the eager warm-up launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stem_wgrad_kernel|stem_fprop_kernel|bn_pool_bwd_apply_kernel|bn_pool_bwd_reduce_pooled_kernel|bn_pool_fwd_kernel|conv_igemm_kernel|conv_wgrad_halo_kernel' -s 30 -c 24 -o gpurun_out/r01c_full_$P python bench.py --precision $P --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
