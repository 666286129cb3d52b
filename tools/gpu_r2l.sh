set -x
mkdir -p gpurun_out
timeout 180 python tests/gpu_diag.py attn_qkv > gpurun_out/r2l_attn_qkv_nomc.log 2>&1; echo "nomc rc=$?"
grep -v PARITY gpurun_out/r2l_attn_qkv_nomc.log | tail -4
ETUDE_DIAG_DEV=1 timeout 180 python tests/gpu_diag.py attn_qkv > gpurun_out/r2l_attn_qkv_split8.log 2>&1; echo "split rc=$?"
grep -v PARITY gpurun_out/r2l_attn_qkv_split8.log | tail -4
timeout 180 python tests/gpu_diag.py attn_qkv_trace > gpurun_out/r2l_attn_qkv_trace_split8.log 2>&1; echo "trace rc=$?"
sed -n 1,3p gpurun_out/r2l_attn_qkv_trace_split8.log; sed -n '/iteration 12/,/iteration 14/p' gpurun_out/r2l_attn_qkv_trace_split8.log
