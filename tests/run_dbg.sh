#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1 $BENCH_ARGS"; env $1 ETUDE_SYNC_DEBUG=1 timeout 300 python bench.py --songs-per-gpu ${SONGS:-2} --steps 1 --warmup 1 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err; echo "[exit $?]"; grep -E "EtudeError|Error" gpurun_out/bench_dbg.err | tail -2; tail -c 100 gpurun_out/bench_dbg.json; echo; }
for i in 1 2 3; do BENCH_ARGS="--window-batch 8" run ETUDE_X=$i; done; for i in 5 6; do BENCH_ARGS="--window-batch 32" run ETUDE_X=$i; done
