set -x
mkdir -p gpurun_out
for i in 1 2; do CHAIN_BIG=1 timeout 300 python tests/gpu_diag.py chain 2>&1 | grep -v PARITY | tail -12; done
timeout 300 python tests/gpu_diag.py chain_trace > gpurun_out/r2n_chain_trace.log 2>&1; echo "trace rc=$?"
head -3 gpurun_out/r2n_chain_trace.log
