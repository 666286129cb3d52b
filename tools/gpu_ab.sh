mkdir -p gpurun_out
for v in $VARIANTS; do
  cp ab/lib_$v.so etude_b200/libetude_b200.so; cp ab/lib_${v}_dev.so etude_b200/libetude_b200_dev.so
  echo "== $v: $(timeout 200 python tests/gpu_diag.py attn_qkv 2>&1 | grep 'S=16384\|FAIL' | head -3)"
  echo "== $v: $(timeout 200 python tests/gpu_diag.py attn_qkv_cross 2>&1 | grep 'PASS\|FAIL' | head -3)"
  timeout 600 python bench.py --songs 32 --no-cpu-baseline --steps 2 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
d = json.load(open('gpurun_out/ab_$v.json'))
print('== $v', round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],1), round(v.get('tflops',0))) for k, v in d['kernels'].items() if k in ('attention','attention_fused','chain','gemm_bias')})
PY
done
