#!/bin/bash
# Quick GPU call: attention diag (+ timeline), pytest -m gpu, one short bench.
mkdir -p gpurun_out
timeout 300 python tests/gpu_diag.py attn > gpurun_out/attn_v4.log 2>&1; echo "[attn exit $?]"; grep -a "ATTN\|rror\|PASS\|FAIL" gpurun_out/attn_v4.log | head -20
if [ -n "$TRACE" ]; then timeout 100 python tests/gpu_diag.py attn_trace 2>&1 | tee gpurun_out/attn_trace.log | sed -n 1,16p; fi
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --songs-per-gpu ${SONGS:-8} --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[bench exit $?]"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"],"clk",d["clocks"])
for k,v in d["kernels"].items():
    print("  %-12s %6.1f launches %8.2f ms/step  share %.3f  %s"%(k,v["launches_per_step"],v["ms_per_step"],v["share_of_step"], ("%.0f TF"%v["tflops"]) if "tflops" in v else ("%.0f GB/s"%v.get("gbs",0))))
PY
tail -3 gpurun_out/bench.err
