#!/bin/bash
# ncu full captures (source-level stall sampling) of the attention kernels at the encoder shape
mkdir -p gpurun_out
B="python bench.py --songs-per-gpu 1 --window-batch 8 --steps 1 --warmup 1 --no-cpu-baseline"
cap() {  # name regex skip count env
  env $5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o /tmp/prof_$1 $B > gpurun_out/ncu_$1.log 2>&1
  echo "[ncu $1 exit $?]"
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv --print-source sass > gpurun_out/$1_source.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  sz=$(stat -c %s /tmp/prof_$1.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 0 ] && [ "$sz" -lt 12000000 ]; then cp /tmp/prof_$1.ncu-rep gpurun_out/; fi
}
cap attn3_enc attention3 11 1 ETUDE_ATTN_V3=1
cap attn2_enc attention2 11 1 ETUDE_X=1
du -sh gpurun_out
