#!/bin/bash
# One GPU call: pytest -m gpu, a short bench, the ncu launch list and full captures of the top kernels.
mkdir -p gpurun_out
SONGS=${SONGS:-8}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "[pytest exit $?]" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --songs-per-gpu $SONGS --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[bench exit $?]"
tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ -n "$NCU" ]; then
B="python bench.py --songs-per-gpu 1 --window-batch 8 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "[ncu list exit $?]"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 11 -c 4 -f -o gpurun_out/prof_chain $B > gpurun_out/ncu_chain.log 2>&1
echo "[ncu chain exit $?]"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention2 -s 11 -c 11 -f -o gpurun_out/prof_attn $B > gpurun_out/ncu_attn.log 2>&1
echo "[ncu attn exit $?]"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 10 -c 5 -f -o gpurun_out/prof_gemm $B > gpurun_out/ncu_gemm.log 2>&1
echo "[ncu gemm exit $?]"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"logmel_kernel|embed_kernel" -s 2 -c 2 -f -o gpurun_out/prof_front $B > gpurun_out/ncu_front.log 2>&1
echo "[ncu front exit $?]"
fi
