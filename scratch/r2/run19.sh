timeout 100 python scratch/r2/rows_time.py 2>&1 | tail -4
