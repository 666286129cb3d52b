for v in cur poly0 old; do
  cp ab/lib_$v.so etude_b200/libetude_b200.so; cp ab/lib_${v}_dev.so etude_b200/libetude_b200_dev.so
  echo "== $v"; timeout 300 python tests/model_flake_diag.py 60 2>&1 | tail -12
done
