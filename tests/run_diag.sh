#!/bin/bash
# Runs every diagnostic stage in its own process with a timeout; collects output under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/diag.log 2>&1
for s in "$@"; do
  timeout 300 python tests/gpu_diag.py $s >> gpurun_out/diag.log 2>&1
  echo "[stage $s exit $?]" >> gpurun_out/diag.log
done
tail -150 gpurun_out/diag.log
