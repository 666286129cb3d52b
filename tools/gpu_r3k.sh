for v in old poly0 cur; do
  cp ab/lib_$v.so etude_b200/libetude_b200.so; cp ab/lib_${v}_dev.so etude_b200/libetude_b200_dev.so
  echo "== $v: $(timeout 200 python tests/dump_encoder.py $v 2>&1 | tail -1)"
done
python tests/dump_encoder.py old poly0; python tests/dump_encoder.py old cur; python tests/dump_encoder.py poly0 cur
rm -f gpurun_out/enc_*.npy
cp ab/lib_cur.so etude_b200/libetude_b200.so; cp ab/lib_cur_dev.so etude_b200/libetude_b200_dev.so
timeout 300 python tests/gpu_diag.py attn_qkv_cross 2>&1 | grep -v PARITY | tail -8
