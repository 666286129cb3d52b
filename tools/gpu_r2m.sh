set -x
mkdir -p gpurun_out
timeout 300 python tests/gpu_diag.py chain > gpurun_out/r2m_chain.log 2>&1; echo "chain rc=$?"
grep -v PARITY gpurun_out/r2m_chain.log | tail -12
timeout 300 python tests/gpu_diag.py chain_trace > gpurun_out/r2m_chain_trace.log 2>&1; echo "trace rc=$?"
head -3 gpurun_out/r2m_chain_trace.log
CHAIN_BIG=1 timeout 300 python tests/gpu_diag.py chain 2>&1 | grep -v PARITY | tail -4
