#!/bin/bash
mkdir -p gpurun_out
for s in ${STAGES:-chain chain_trace}; do timeout 300 python tests/gpu_diag.py $s > gpurun_out/diag_$s.log 2>&1; echo "[stage $s exit $?]"; grep -v "  TMA  " gpurun_out/diag_$s.log | tail -${TAILN:-60}; done
if [ -n "$FULL" ]; then bash tests/run_quick.sh; fi
