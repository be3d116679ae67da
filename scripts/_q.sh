timeout 600 python -m pytest tests/test_gpu_extensions.py -q -k p3m_influence 2>&1 | grep -E "FAILED|passed|failed"
