#!/bin/bash
# One GPU call: pytest -m gpu, a short bench, the ncu launch list and full captures of the top kernels.
mkdir -p gpurun_out
SONGS=${SONGS:-8}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --songs-per-gpu $SONGS --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[bench exit $?]"
tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ -n "$NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --songs-per-gpu 1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "[ncu list exit $?]"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 6 -c 2 -f -o gpurun_out/prof_attn \
    python bench.py --songs-per-gpu 1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
echo "[ncu attn exit $?]"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 10 -c 4 -f -o gpurun_out/prof_gemm \
    python bench.py --songs-per-gpu 1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
echo "[ncu gemm exit $?]"
fi
