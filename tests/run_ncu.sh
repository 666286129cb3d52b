#!/bin/bash
# ncu captures of the top kernels; .ncu-rep files stay in /tmp on the box (too big), text exports come back.
mkdir -p gpurun_out
B="python bench.py --songs-per-gpu 1 --window-batch 8 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "[ncu list exit $?]"
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o /tmp/prof_$1 $B > gpurun_out/ncu_$1.log 2>&1
  echo "[ncu $1 exit $?]"
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv --print-source sass > gpurun_out/$1_source.csv 2>/dev/null
  ls -la /tmp/prof_$1.ncu-rep
  sz=$(stat -c %s /tmp/prof_$1.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 0 ] && [ "$sz" -lt 9000000 ]; then cp /tmp/prof_$1.ncu-rep gpurun_out/; fi
}
for k in ${KERNELS:-chain attn_enc attn_cross attn_self88 attn_time gemm front}; do
case $k in
  chain) cap chain chain_kernel 11 1;;
  attn_enc) cap attn_enc attention2 11 1;;
  attn_cross) cap attn_cross attention2 14 1;;
  attn_self88) cap attn_self88 attention2 15 1;;
  attn_time) cap attn_time attention2 19 1;;
  gemm) cap gemm gemm_tcgen05 10 1;;
  front) cap front "logmel_kernel|embed_kernel" 2 2;;
esac
done
du -sh gpurun_out
