python -m pytest tests/test_gpu_assets.py tests/test_gpu_abi.py -m gpu -x -q 2>&1 | tail -15
