timeout 600 python -m pytest tests/test_slope_gpu.py tests/test_forms_gpu.py -m gpu -q 2>&1 | tail -6
timeout 300 python examples/slope_stability.py 25 25 2>&1 | tail -6
