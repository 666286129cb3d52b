#!/bin/bash
# One GPU call: pytest -m gpu, smoke, attention3 vs attention2 diag, short benches, the ncu launch list.
mkdir -p gpurun_out
SONGS=${SONGS:-8}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tests/gpu_diag.py attn > gpurun_out/attn_v2.log 2>&1; echo "[attn v2 exit $?]"; grep ATTN gpurun_out/attn_v2.log
ETUDE_ATTN_V3=1 timeout 300 python tests/gpu_diag.py attn > gpurun_out/attn_v3.log 2>&1; V3=$?; echo "[attn v3 exit $V3]"; grep -a "ATTN\|rror" gpurun_out/attn_v3.log | head -20
ETUDE_E2E_TRACE=1 timeout 900 python bench.py --songs-per-gpu $SONGS --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[bench exit $?]"
tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$V3" = "0" ]; then
ETUDE_ATTN_V3=1 timeout 900 python bench.py --songs-per-gpu $SONGS --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; echo "[bench v3 exit $?]"
tail -c 6000 gpurun_out/bench_v3.json; tail -5 gpurun_out/bench_v3.err
ETUDE_ATTN_V3=1 timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_v3.log 2>&1; echo "[pytest v3 exit $?]"; tail -5 gpurun_out/pytest_gpu_v3.log
fi
if [ -n "$NCU" ]; then
B="python bench.py --songs-per-gpu 1 --window-batch 8 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "[ncu list exit $?]"
fi
