#!/bin/bash
# ncu --set full capture of one FFN chain launch at the bench's window batch (32): DRAM traffic per launch for bench.py's
# roofline.traffic (profiles/chain_traffic.json is derived from it).
mkdir -p gpurun_out
B="python bench.py --songs-per-gpu 2 --window-batch 32 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain2_kernel -s 2 -c 1 -f -o /tmp/prof_chain $B > gpurun_out/ncu_chain.log 2>&1
echo "[ncu chain exit $?]"
ncu -i /tmp/prof_chain.ncu-rep --page details > gpurun_out/chain_full_details.txt 2>/dev/null
ncu -i /tmp/prof_chain.ncu-rep --page raw --csv > gpurun_out/chain_full_raw.csv 2>/dev/null
grep -E "chain2_kernel|Duration|DRAM Throughput" gpurun_out/chain_full_details.txt | head -5
