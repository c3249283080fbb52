# final round-2 evidence on one GPU, last build (bn_pool index arithmetic / occupancy on top of run_final3.sh's build)
O=gpurun_out/final4
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 --kernel-table $O/kernel_table_bf16.json > $O/bench_bf16_n1.json 2> $O/bench_bf16.err; cut -c1-200 $O/bench_bf16_n1.json
for P in tf32 tf32x3; do
  timeout 300 python bench.py --precision $P --steps 20 --warmup 5 --no-cpu-baseline --kernel-table $O/kernel_table_$P.json > $O/bench_${P}_n1.json 2> $O/bench_$P.err; cut -c1-120 $O/bench_${P}_n1.json
done
timeout 300 python bench.py --config cfg4 --steps 20 --warmup 5 --no-cpu-baseline --kernel-table $O/kernel_table_cfg4.json > $O/bench_cfg4.json 2> $O/bench_cfg4.err; cut -c1-120 $O/bench_cfg4.json
timeout 300 python bench.py --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline > $O/plain_ll.json 2> $O/plain_ll.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bf16.csv python bench.py --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_ll.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stem_wgrad_kernel|stem_rows_fprop_kernel|bn_pool_bwd_apply_kernel|bn_pool_bwd_reduce_pooled_kernel|bn_pool_fwd_kernel|conv_igemm_kernel|conv_rows_igemm_kernel|conv_wgrad_halo_kernel' -s 30 -c 26 -o $O/full_bf16 -f python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
ls $O | head -40
