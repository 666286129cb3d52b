#!/bin/bash
mkdir -p gpurun_out
for s in ${STAGES:-chain_trace notes}; do timeout 300 python tests/gpu_diag.py $s > gpurun_out/diag_$s.log 2>&1; echo "[stage $s exit $?]"; tail -${TAILN:-150} gpurun_out/diag_$s.log; done
