O=gpurun_out/ncu2
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stem_wgrad_kernel|bn_pool_bwd_apply' -s 8 -c 6 -o $O/stem_wgrad -f python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_stem.log 2>&1; tail -3 $O/ncu_stem.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lstm_seq' -s 4 -c 2 -o $O/lstm -f python bench.py --config cfg5 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_lstm.log 2>&1; tail -3 $O/ncu_lstm.log
ls -la $O
