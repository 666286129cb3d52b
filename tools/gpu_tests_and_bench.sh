mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/rT_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/rT_pytest_gpu.log
cp gpurun_out/parity_margins.txt gpurun_out/rT_parity_margins.txt 2>/dev/null
timeout 600 python bench.py --config 3 --no-cpu-baseline > gpurun_out/rT_bench_config3.json 2> gpurun_out/rT_bench_config3.err; echo "bench3 rc=$?"
timeout 900 python bench.py > gpurun_out/rT_bench_default.json 2> gpurun_out/rT_bench_default.err; echo "bench default rc=$?"
python - <<PY
import json
for f in ('rT_bench_config3', 'rT_bench_default'):
    d = json.load(open(f'gpurun_out/{f}.json'))
    print(f, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'], 1), d['clocks']['sm_mhz'], d['roofline']['kernel'], round(d['roofline']['achieved']), round(d['roofline']['frac'], 3),
          {k: (round(v['ms_per_step'],1), round(v.get('tflops',0))) for k, v in d['kernels'].items() if k in ('attention','attention_fused','chain','gemm_bias')})
PY
