timeout 600 python -m pytest tests/test_gpu_tf32.py -x -q 2>&1 | tail -15
