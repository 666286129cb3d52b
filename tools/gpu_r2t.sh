mkdir -p gpurun_out
cp ab/lib_poly.so etude_b200/libetude_b200.so; cp ab/lib_poly_dev.so etude_b200/libetude_b200_dev.so
timeout 600 ncu --set full --clock-control none -k regex:attention4 -s 2 -c 1 -f -o gpurun_out/r2t_attn4_poly python tests/gpu_diag.py attn_trace > gpurun_out/r2t_ncu_poly.log 2>&1; echo "ncu rc=$?"
cp ab/lib_pivot.so etude_b200/libetude_b200.so; cp ab/lib_pivot_dev.so etude_b200/libetude_b200_dev.so
timeout 600 ncu --set full --clock-control none -k regex:attention4 -s 2 -c 1 -f -o gpurun_out/r2t_attn4_base python tests/gpu_diag.py attn_trace > gpurun_out/r2t_ncu_base.log 2>&1; echo "ncu rc=$?"
