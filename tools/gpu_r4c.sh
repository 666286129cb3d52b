timeout 180 python tests/gpu_diag.py attn_qkv_trace > gpurun_out/r4c_attn_pair_trace.log 2>&1; echo "trace rc=$?"
sed -n 1,3p gpurun_out/r4c_attn_pair_trace.log; sed -n '/iteration 12/,/iteration 15/p' gpurun_out/r4c_attn_pair_trace.log | grep -v "\.1 "
