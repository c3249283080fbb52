timeout 120 python scratch/stem_prof.py 2>&1 | tail -6
bash scratch/run1.sh > /dev/null 2>&1
tail -3 gpurun_out/t.log
