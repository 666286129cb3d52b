mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2x_pytest_gpu.log
cp gpurun_out/parity_margins.txt gpurun_out/r2x_parity_margins.txt 2>/dev/null
timeout 600 python bench.py --config 3 --no-cpu-baseline > gpurun_out/r2x_bench_config3.json 2> gpurun_out/r2x_bench_config3.err; echo "bench3 rc=$?"
python - <<PY
import json
d = json.load(open('gpurun_out/r2x_bench_config3.json'))
print('config3', round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['clocks']['sm_mhz'])
PY
