mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --no-cpu-baseline > gpurun_out/rN_bench_n$N.json 2> gpurun_out/rN_bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
txt = [l for l in open('gpurun_out/rN_bench_n$N.json') if l.startswith('{')][-1]
d = json.loads(txt)
print('n$N', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'], 1), d['n_gpus'], d['scaling'], d['clocks']['sm_mhz'])
PY
