mkdir -p gpurun_out
timeout 200 python tests/gpu_diag.py attn_qkv 2>&1 | grep -v PARITY | tail -9
timeout 180 python tests/gpu_diag.py attn_qkv_trace > gpurun_out/r3d_attn_pair_trace.log 2>&1; echo "trace rc=$?"
sed -n 1,3p gpurun_out/r3d_attn_pair_trace.log; sed -n '/iteration 12/,/iteration 15/p' gpurun_out/r3d_attn_pair_trace.log
timeout 600 python bench.py --songs 32 --no-cpu-baseline --steps 2 > gpurun_out/r3d_bench32.json 2> gpurun_out/r3d_bench32.err
python - <<PY
import json
d = json.load(open('gpurun_out/r3d_bench32.json'))
print('== pair', round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'], {k: (round(v['ms_per_step'],1), round(v.get('tflops',0))) for k, v in d['kernels'].items() if k in ('attention','attention_fused','chain','gemm_bias')})
PY
