#!/bin/bash
mkdir -p gpurun_out
which compute-sanitizer || ls /usr/local/cuda/bin | grep -i sanit
ETUDE_SYNC_DEBUG=1 timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python bench.py --songs-per-gpu 1 --window-batch 8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/san.log 2>&1; echo "[san exit $?]"
grep -v "^$" gpurun_out/san.log | head -60
