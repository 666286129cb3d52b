#!/bin/bash
# attention v2 bring-up: kernel diagnostics (each stage in its own process), then the GPU test suite and a short bench
mkdir -p gpurun_out
: > gpurun_out/diag.log
for s in chain attn gemm model; do timeout 300 python tests/gpu_diag.py $s >> gpurun_out/diag.log 2>&1; echo "[stage $s exit $?]" >> gpurun_out/diag.log; done
tail -70 gpurun_out/diag.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --songs-per-gpu 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[bench exit $?]"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'util',d['model'])
    for k,v in d['kernels'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
except Exception as e: print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-3000:])
PY
