mkdir -p gpurun_out
timeout 300 python tests/gpu_diag.py attn_qkv 2>&1 | grep -v PARITY | tail -30
