set -x
mkdir -p gpurun_out
timeout 180 python tests/gpu_diag.py attn_qkv > gpurun_out/r2e_attn_qkv_push.log 2>&1; echo "push rc=$?"
grep -v PARITY gpurun_out/r2e_attn_qkv_push.log | tail -9
ETUDE_DIAG_DEV=1 timeout 180 python tests/gpu_diag.py attn_qkv > gpurun_out/r2e_attn_qkv_pull.log 2>&1; echo "pull rc=$?"
grep -v PARITY gpurun_out/r2e_attn_qkv_pull.log | tail -5
timeout 180 python tests/gpu_diag.py attn_qkv_trace > gpurun_out/r2e_attn_qkv_trace_pull.log 2>&1; echo "trace rc=$?"
sed -n 1,3p gpurun_out/r2e_attn_qkv_trace_pull.log; sed -n '/iteration 12/,/iteration 15/p' gpurun_out/r2e_attn_qkv_trace_pull.log
