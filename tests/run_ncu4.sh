#!/bin/bash
# r1f: chain timeline + ncu full captures of the current top kernels (details + raw pages only)
mkdir -p gpurun_out
timeout 120 python tests/gpu_diag.py chain_trace > gpurun_out/chain_trace.log 2>&1; echo "[chain_trace exit $?]"
B="python bench.py --songs-per-gpu 1 --window-batch 16 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "[ncu list exit $?]"
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o /tmp/prof_$1 $B > gpurun_out/ncu_$1.log 2>&1
  echo "[ncu $1 exit $?]"
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
}
cap gemm_bstat gemm_bstat 6 1
cap chain2 chain2_kernel 6 1
cap attn4 attention4 6 1
cap embed2 embed2_kernel 1 1
cap logmel2 logmel2_kernel 1 1
du -sh gpurun_out
