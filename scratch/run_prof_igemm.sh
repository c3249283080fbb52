timeout 120 python scratch/prof_igemm.py 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_igemm' -s 46 -c 3 -o gpurun_out/r01e_igemm python scratch/prof_igemm.py > gpurun_out/ncu_igemm.log 2>&1
