timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "10k or gradient_method_fp32" 2>&1 | tail -12
