mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tf32.py -q -k "pair or tf32_vs_oracle or bf16_vs_oracle or exact" 2>&1 | tail -5
SN_IGEMM_PAIR=1 timeout 60 python scratch/pair_check.py > gpurun_out/pair1.log 2>&1; echo "pair1 exit $?"; tail -7 gpurun_out/pair1.log
