set -x
mkdir -p gpurun_out
timeout 300 python bench.py --config 2 > gpurun_out/r2j_frontend_staged.json 2> gpurun_out/r2j_frontend_staged.err; echo "staged rc=$?"
timeout 300 python tests/gpu_diag.py logmel > gpurun_out/r2j_logmel_staged.log 2>&1; echo "logmel diag rc=$?"; grep -v PARITY gpurun_out/r2j_logmel_staged.log | tail -7
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "logmel or wav2feature or smoke" > gpurun_out/r2j_pytest_logmel.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2j_pytest_logmel.log
ETUDE_NVCC_FLAGS=-DETUDE_LOGMEL_STAGED=0 python etude_b200/build.py --force > gpurun_out/r2j_rebuild.log 2>&1; echo "rebuild rc=$?"
timeout 300 python bench.py --config 2 > gpurun_out/r2j_frontend_unstaged.json 2> gpurun_out/r2j_frontend_unstaged.err; echo "unstaged rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2j_frontend_staged.json', 'gpurun_out/r2j_frontend_unstaged.json'):
    try:
        d = json.load(open(f)); print(f, round(d['value']), 'audio-s/s', round(d['ms_per_step'], 3), 'ms', round(d['roofline']['achieved'], 1), 'GB/s', round(d['roofline']['frac'], 4), d['clocks'])
    except Exception as e: print('no json', f, e)
PY
