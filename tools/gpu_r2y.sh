mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logmel2 -s 3 -c 1 -f -o gpurun_out/r2y_logmel2 python bench.py --config 2 --steps 3 > gpurun_out/r2y_ncu_logmel2.log 2>&1; echo "ncu logmel2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_qkv -s 8 -c 1 -f -o gpurun_out/r2y_attn_qkv python tests/gpu_diag.py attn_qkv > gpurun_out/r2y_ncu_attn_qkv.log 2>&1; echo "ncu attn_qkv rc=$?"
ls -la gpurun_out/*.ncu-rep
