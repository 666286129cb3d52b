#!/bin/bash
# pytest -m gpu + a short bench (with the e2e wall-time breakdown on stderr)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]"; tail -15 gpurun_out/pytest_gpu.log
ETUDE_E2E_TRACE=1 timeout 900 python bench.py --songs-per-gpu ${SONGS:-8} --steps 3 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[bench exit $?]"
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'util',d['model'])
    for k,v in d['kernels'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
except Exception as e: print('bench parse failed', e)
PY
