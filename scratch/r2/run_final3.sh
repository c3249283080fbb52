# final round-2 evidence on one GPU (the build after the TMA reduce-add epilogues, bf16 stem wgrad, tensor-pipe LSTM)
O=gpurun_out/final3
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 --kernel-table $O/kernel_table_bf16.json > $O/bench_bf16_n1.json 2> $O/bench_bf16.err; cut -c1-200 $O/bench_bf16_n1.json
for P in tf32 tf32x3 fp32; do
  timeout 300 python bench.py --precision $P --steps 20 --warmup 5 --no-cpu-baseline --kernel-table $O/kernel_table_$P.json > $O/bench_${P}_n1.json 2> $O/bench_$P.err; cut -c1-120 $O/bench_${P}_n1.json
done
for C in cfg1 cfg2 cfg4 cfg5; do
  timeout 300 python bench.py --config $C --steps 20 --warmup 5 --no-cpu-baseline --kernel-table $O/kernel_table_$C.json > $O/bench_$C.json 2> $O/bench_$C.err; cut -c1-120 $O/bench_$C.json
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_ref.err; cut -c1-160 $O/bench_reference.json
timeout 300 python bench.py --micro > $O/micro.json 2> $O/micro.err; tail -3 $O/micro.err
timeout 300 python bench.py --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline > $O/plain_ll.json 2> $O/plain_ll.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bf16.csv python bench.py --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_ll.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stem_wgrad_kernel|stem_rows_fprop_kernel|bn_pool_bwd_apply_kernel|bn_pool_bwd_reduce_pooled_kernel|bn_pool_fwd_kernel|conv_igemm_kernel|conv_rows_igemm_kernel|conv_wgrad_halo_kernel' -s 30 -c 26 -o $O/full_bf16 -f python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'lstm_seq' -s 4 -c 2 -o $O/full_cfg5 -f python bench.py --config cfg5 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_cfg5.log 2>&1
ls -la $O | head -50
