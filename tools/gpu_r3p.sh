mkdir -p gpurun_out
# launch list of one small step (per-launch times are cold-cache and serialised: the SHARE per kernel is what is compared)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r3p_launches.csv python bench.py --songs 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r3p_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_pair -s 8 -c 1 -f -o gpurun_out/r3p_attn_pair python tests/gpu_diag.py attn_qkv > gpurun_out/r3p_ncu_attn_pair.log 2>&1; echo "ncu attn_pair rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain3 -s 3 -c 1 -f -o gpurun_out/r3p_chain3 python tests/gpu_diag.py chain_trace > gpurun_out/r3p_ncu_chain3.log 2>&1; echo "ncu chain3 rc=$?"
ls -la gpurun_out/r3p_*
