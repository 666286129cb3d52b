timeout 200 python tests/gpu_diag.py pairmma_bench 2>&1 | tail -12
