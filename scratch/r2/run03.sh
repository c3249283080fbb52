timeout 600 python scratch/r2/grad_diag.py 0.1 2>&1 | tail -6
