mkdir -p gpurun_out
timeout 300 python tests/gpu_diag.py chain 2>&1 | grep -v PARITY | tail -10
timeout 300 python tests/gpu_diag.py attn 2>&1 | grep -v PARITY | tail -10
for v in low high low high; do
  cp ab/lib_$v.so etude_b200/libetude_b200.so; cp ab/lib_${v}_dev.so etude_b200/libetude_b200_dev.so
  timeout 600 python bench.py --songs 32 --no-cpu-baseline --steps 2 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
d = json.load(open('gpurun_out/ab_$v.json'))
print('== $v', round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],1), round(v.get('tflops',0))) for k, v in d['kernels'].items() if k in ('attention','attention_fused','chain','gemm_bias')})
PY
done
