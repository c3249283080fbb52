timeout 120 python scratch/stem_prof.py 2>&1 | tail -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stem_' -s 8 -c 4 -o gpurun_out/r01d_stem python scratch/stem_prof.py > gpurun_out/ncu_stem.log 2>&1
