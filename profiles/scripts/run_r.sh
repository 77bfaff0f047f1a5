python -m pytest tests/test_gpu_ref_goldens.py tests/test_gpu_scripts.py tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -40
