timeout 600 python -m pytest tests/test_gpu_data_cache.py tests/test_gpu_step.py -x -q 2>&1 | tail -15
