for v in cur poly0; do
  cp ab/lib_$v.so etude_b200/libetude_b200.so; cp ab/lib_${v}_dev.so etude_b200/libetude_b200_dev.so
  echo "== $v"; timeout 200 python tests/model_frames_diag.py 2>&1 | tail -9
done
