mkdir -p gpurun_out
for v in old poly0 cur; do
  cp ab/lib_$v.so etude_b200/libetude_b200.so; cp ab/lib_${v}_dev.so etude_b200/libetude_b200_dev.so
  echo "== $v"; timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "9tuple or encode_decode" 2>&1 | grep "PARITY\|passed\|failed" | cut -c1-400
done
