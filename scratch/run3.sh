timeout 120 python scratch/stem_prof.py 2>&1 | tail -6
timeout 600 python -m pytest tests/test_gpu_tf32.py -x -q 2>&1 | tail -3
