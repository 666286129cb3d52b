set -x
mkdir -p gpurun_out
timeout 180 python tests/gpu_diag.py attn_qkv_trace > gpurun_out/r2c_attn_qkv_trace.log 2>&1; echo "trace rc=$?"
cat gpurun_out/r2c_attn_qkv_trace.log
