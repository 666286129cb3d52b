timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "outlier" 2>&1 | grep "PARITY\|passed\|failed\|Error\|assert" | cut -c1-600
