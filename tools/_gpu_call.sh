timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_search or gradient" 2>&1 | tail -15
