for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep "^E   assert\|^FAILED\|passed\|failed" | head -6; done
SN_DEBUG_POISON=1 timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep "^E  \|^FAILED\|passed\|failed" | head -6
