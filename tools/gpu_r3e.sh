mkdir -p gpurun_out
timeout 200 python tests/gpu_diag.py attn_qkv 2>&1 | grep -v PARITY | tail -4
timeout 180 python tests/gpu_diag.py attn_qkv_trace > gpurun_out/r3e_attn_pair_trace.log 2>&1; echo "trace rc=$?"
sed -n 1,3p gpurun_out/r3e_attn_pair_trace.log; sed -n '/iteration 12/,/iteration 14/p' gpurun_out/r3e_attn_pair_trace.log
