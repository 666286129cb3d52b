mkdir -p gpurun_out
timeout 120 python tests/gpu_diag.py pairmma 2>&1 | tail -12
