for i in 1 2 3 4; do timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -A12 "^E\|^FAILED\|passed\|failed" | head -30; done
