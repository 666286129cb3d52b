mkdir -p gpurun_out
for wb in 32 64 16 32 64; do
  timeout 600 python bench.py --songs 32 --window-batch $wb --no-cpu-baseline --steps 2 > gpurun_out/r3s_wb$wb.json 2> gpurun_out/r3s_wb$wb.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r3s_wb$wb.json'))
print('== window batch $wb', round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],1), round(v.get('tflops',0))) for k, v in d['kernels'].items() if k in ('attention','attention_fused','chain','gemm_bias')})
PY
done
