timeout 120 python scratch/stem_dbg.py fprop 2>&1 | tail -4
timeout 120 python scratch/stem_dbg.py wgrad 2>&1 | tail -4
