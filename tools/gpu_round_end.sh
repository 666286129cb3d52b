mkdir -p gpurun_out
rm -f gpurun_out/parity_margins.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/rZ_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/rZ_pytest_gpu.log
cp gpurun_out/parity_margins.txt gpurun_out/rZ_parity_margins.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/rZ_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/rZ_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 > gpurun_out/rZ_bench_reference_arm.json 2> gpurun_out/rZ_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 1200 python bench.py > gpurun_out/rZ_bench_default.json 2> gpurun_out/rZ_bench_default.err; echo "bench default rc=$?"
timeout 600 python bench.py --config 3 > gpurun_out/rZ_bench_config3.json 2> gpurun_out/rZ_bench_config3.err; echo "bench config3 rc=$?"
timeout 600 python bench.py --config 2 > gpurun_out/rZ_bench_config2.json 2> gpurun_out/rZ_bench_config2.err; echo "bench config2 rc=$?"
python - <<PY
import json
for f in ('rZ_bench_reference_arm', 'rZ_bench_default', 'rZ_bench_config3', 'rZ_bench_config2'):
    d = json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
    print(f, round(d['value'], 1), d.get('e2e', {}).get('value'), d.get('ms_per_step'), (d.get('clocks') or {}).get('sm_mhz'), (d.get('roofline') or {}).get('frac'), d.get('cpu_baseline'))
PY
