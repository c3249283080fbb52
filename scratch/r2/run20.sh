mkdir -p gpurun_out/r2t
timeout 100 python scratch/r2/rows_one.py && timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv_rows_igemm_kernel|conv_igemm_kernel' -s 2 -c 1 -o gpurun_out/r2t/rows python scratch/r2/rows_one.py > gpurun_out/r2t/ncu_rows.log 2>&1
SN_IGEMM_ROWS=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv_rows_igemm_kernel|conv_igemm_kernel' -s 2 -c 1 -o gpurun_out/r2t/old python scratch/r2/rows_one.py > gpurun_out/r2t/ncu_old.log 2>&1
ls -la gpurun_out/r2t
