set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.txt
timeout 180 python tests/gpu_diag.py attn_qkv > gpurun_out/r2g_attn_qkv.log 2>&1; echo "attn_qkv rc=$?"
grep -v PARITY gpurun_out/r2g_attn_qkv.log | tail -9
timeout 180 python tests/gpu_diag.py attn_qkv_trace > gpurun_out/r2g_attn_qkv_trace.log 2>&1; echo "trace rc=$?"
sed -n 1,3p gpurun_out/r2g_attn_qkv_trace.log; sed -n '/iteration 12/,/iteration 15/p' gpurun_out/r2g_attn_qkv_trace.log
timeout 600 python bench.py --songs 32 --no-cpu-baseline > gpurun_out/r2g_bench_32songs.json 2> gpurun_out/r2g_bench_32songs.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2g_bench_32songs.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2g_bench_32songs.json'))
    print(d['value'], d['e2e']['value'], d['clocks'])
    for k, v in d['kernels'].items(): print('   ', k, round(v['ms_per_step'], 2), 'ms', round(v.get('share_of_kernel_time', 0), 3), round(v.get('tflops', 0)), round(v.get('gbs', 0)))
except Exception as e: print('no bench json', e)
PY
timeout 900 python -m pytest tests -m gpu -x -q -rP > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2g_pytest_gpu.log
