mkdir -p gpurun_out; rm -f gpurun_out/parity_margins.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r4f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r4f_pytest_gpu.log
cp gpurun_out/parity_margins.txt gpurun_out/r4f_parity_margins.txt
