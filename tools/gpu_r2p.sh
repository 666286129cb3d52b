set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.txt
timeout 900 python -m pytest tests -m gpu -x -q -rP > gpurun_out/r2p_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2p_pytest_gpu.log
timeout 600 python bench.py --songs 32 --no-cpu-baseline > gpurun_out/r2p_bench_32songs.json 2> gpurun_out/r2p_bench_32songs.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2p_bench_32songs.err
timeout 600 python bench.py --config 3 --no-cpu-baseline > gpurun_out/r2p_bench_config3.json 2> gpurun_out/r2p_bench_config3.err; echo "bench3 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2p_bench_32songs.json','gpurun_out/r2p_bench_config3.json'):
  try:
    d = json.load(open(f))
    print(f, d['value'], d['e2e']['value'], d['clocks'])
    for k, v in d['kernels'].items(): print('   ', k, round(v['ms_per_step'], 2), 'ms', round(v.get('share_of_kernel_time', 0), 3), round(v.get('tflops', 0)), round(v.get('gbs', 0)))
  except Exception as e: print('no bench json', e)
PY
