set -x
mkdir -p gpurun_out
CHAIN_BIG=1 timeout 300 python tests/gpu_diag.py chain 2>&1 | grep -v PARITY | tail -13
timeout 300 python tests/gpu_diag.py chain_trace > gpurun_out/r2r_chain_trace.log 2>&1; echo "trace rc=$?"
head -3 gpurun_out/r2r_chain_trace.log
awk '/tile 6/ && /EPI|MMA  tile 6 ev (0|1|2|3|43)$/' gpurun_out/r2r_chain_trace.log; grep "EPI  tile 7 ev 0" gpurun_out/r2r_chain_trace.log
timeout 600 python bench.py --songs 32 --no-cpu-baseline > gpurun_out/r2r_bench_32songs.json 2> gpurun_out/r2r_bench_32songs.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2r_bench_32songs.json',):
  try:
    d = json.load(open(f))
    print(f, d['value'], d['e2e']['value'], d['clocks'])
    for k, v in d['kernels'].items(): print('   ', k, round(v['ms_per_step'], 2), 'ms', round(v.get('share_of_kernel_time', 0), 3), round(v.get('tflops', 0)), round(v.get('gbs', 0)))
  except Exception as e: print('no bench json', e)
PY
