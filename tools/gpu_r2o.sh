set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain3 -s 3 -c 1 -f -o gpurun_out/r2o_chain3 python tests/gpu_diag.py chain_trace > gpurun_out/r2o_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2o_ncu.log
ls -la gpurun_out/
