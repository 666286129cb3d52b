#!/bin/bash
# Round-end verification on one GPU: pytest -m gpu, smoke, default bench (+ cpu baseline), reference arm, ncu launch list +
# full capture of the attention kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "[smoke exit $?]"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "[default bench exit $?]"
timeout 900 python bench.py --songs-per-gpu 8 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[8-song bench exit $?]"
python - <<'PY'
import json
for f in ("gpurun_out/bench_default.json", "gpurun_out/bench.json"):
    d=json.load(open(f))
    print(f, "value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],1),"clk",d["clocks"]["sm_mhz"],"roofline",d["roofline"]["kernel"],round(d["roofline"]["frac"],3),"util",round(d["model"]["tensor_util_of_sustained_peak"],3),"frontend",round(d["frontend"]["frac"],4),d.get("cpu_baseline",{}).get("value"))
    for k,v in d["kernels"].items():
        print("  %-12s %6.1f launches %8.2f ms/step  share %.3f  %s"%(k,v["launches_per_step"],v["ms_per_step"],v["share_of_step"], ("%.0f TF"%v["tflops"]) if "tflops" in v else ("%.0f GB/s"%v.get("gbs",0))))
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2>&1; echo "[ref exit $?]"
B="python bench.py --songs-per-gpu 1 --window-batch 16 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "[ncu list exit $?]"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention4 -s 6 -c 1 -f -o /tmp/prof_attn4 $B > gpurun_out/ncu_attn4.log 2>&1
echo "[ncu attn4 exit $?]"
ncu -i /tmp/prof_attn4.ncu-rep --page details > gpurun_out/attn4_details.txt 2>/dev/null
ncu -i /tmp/prof_attn4.ncu-rep --page raw --csv > gpurun_out/attn4_raw.csv 2>/dev/null
