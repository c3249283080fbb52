mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_recurrent.py -q 2>&1 | tail -25
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cfg5.csv python bench.py --config cfg5 --precision fp32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu5.log 2>&1; echo ncu exit $?
