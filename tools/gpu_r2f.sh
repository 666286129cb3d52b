set -x
mkdir -p gpurun_out
timeout 180 python tests/gpu_diag.py attn_qkv_trace > gpurun_out/r2f_attn_qkv_trace_push.log 2>&1; echo "trace rc=$?"
sed -n 1,3p gpurun_out/r2f_attn_qkv_trace_push.log; sed -n '/iteration 11/,/iteration 15/p' gpurun_out/r2f_attn_qkv_trace_push.log
