mkdir -p gpurun_out
for v in cur poly0; do
  cp ab/lib_$v.so etude_b200/libetude_b200.so; cp ab/lib_${v}_dev.so etude_b200/libetude_b200_dev.so
  echo "== $v"; timeout 300 python tests/gpu_diag.py attn_qkv 2>&1 | grep -v PARITY | grep -A8 "x\*24"
done
