python -m pytest tests/test_forms_gpu.py tests/test_cylinder_gpu.py -m gpu -q 2>&1 | tail -5
