mkdir -p gpurun_out
timeout 300 python tests/gpu_diag.py attn 2>&1 | grep -v PARITY | tail -10
timeout 300 python tests/gpu_diag.py attn_trace > gpurun_out/r2s_attn_trace.log 2>&1
sed -n 1,12p gpurun_out/r2s_attn_trace.log
