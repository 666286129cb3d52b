mkdir -p gpurun_out
for rep in 1 2 3; do
for v in pivot trim; do
  cp ab/lib_$v.so etude_b200/libetude_b200.so; cp ab/lib_${v}_dev.so etude_b200/libetude_b200_dev.so
  echo "== $v rep $rep: $(timeout 300 python tests/gpu_diag.py chain_trace 2>&1 | grep CHAIN_TRACE)"
done; done
