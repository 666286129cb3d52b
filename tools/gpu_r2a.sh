set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.txt
python -m pytest tests -m gpu -x -q -rP > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2a_pytest_gpu.log
python bench.py --config 3 --no-cpu-baseline > gpurun_out/r2a_bench_config3.json 2> gpurun_out/r2a_bench_config3.err; echo "bench3 rc=$?"
python bench.py > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err; echo "bench rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention4_kernel -s 2 -c 1 -o gpurun_out/r2a_attn4_128 python tools/run_attn_shape.py > gpurun_out/r2a_ncu_attn.log 2>&1; echo "ncu rc=$?"
tail -c 1500 gpurun_out/r2a_bench_config3.err gpurun_out/r2a_bench_default.err
cut -c1-900 gpurun_out/r2a_bench_config3.json gpurun_out/r2a_bench_default.json
